"""flow_log.proto (message/flow_log.proto:14-310) as python-protobuf message classes,
built from a descriptor assembled in code (no protoc in the image).  Used to
cross-check the synthetic generator's encoder and the oracle's decoder against
Google's protobuf implementation."""
from google.protobuf import descriptor_pb2, descriptor_pool, message_factory

T = descriptor_pb2.FieldDescriptorProto
U32, U64, I32, STR, BYT, MSG, DBL = (T.TYPE_UINT32, T.TYPE_UINT64, T.TYPE_INT32, T.TYPE_STRING, T.TYPE_BYTES,
                                     T.TYPE_MESSAGE, T.TYPE_DOUBLE)

SCHEMA = {
    "FlowKey": [("vtap_id", 1, U32), ("tap_type", 2, U32), ("tap_port", 3, U64), ("mac_src", 4, U64),
                ("mac_dst", 5, U64), ("ip_src", 6, U32), ("ip_dst", 7, U32), ("ip6_src", 8, BYT),
                ("ip6_dst", 9, BYT), ("port_src", 10, U32), ("port_dst", 11, U32), ("proto", 12, U32)],
    "FlowMetricsPeer": [("byte_count", 1, U64), ("l3_byte_count", 2, U64), ("l4_byte_count", 3, U64),
                        ("packet_count", 4, U64), ("total_byte_count", 5, U64), ("total_packet_count", 6, U64),
                        ("first", 7, U64), ("last", 8, U64), ("tcp_flags", 9, U32), ("l3_epc_id", 10, I32),
                        ("is_l2_end", 11, U32), ("is_l3_end", 12, U32), ("is_active_host", 13, U32),
                        ("is_device", 14, U32), ("is_vip_interface", 15, U32), ("is_vip", 16, U32),
                        ("real_ip", 20, U32), ("real_port", 21, U32), ("gpid", 22, U32)],
    "TunnelField": [("tx_ip0", 1, U32), ("tx_ip1", 2, U32), ("rx_ip0", 3, U32), ("rx_ip1", 4, U32),
                    ("tx_mac0", 5, U32), ("tx_mac1", 6, U32), ("rx_mac0", 7, U32), ("rx_mac1", 8, U32),
                    ("tx_id", 9, U32), ("rx_id", 10, U32), ("tunnel_type", 11, U32), ("tier", 12, U32),
                    ("is_ipv6", 13, U32)],
    "TcpPerfCountsPeer": [("retrans_count", 1, U32), ("zero_win_count", 2, U32), ("ooo_count", 3, U32)],
    "TCPPerfStats": [("rtt_client_max", 1, U32), ("rtt_server_max", 2, U32), ("srt_max", 3, U32),
                     ("art_max", 4, U32), ("rtt", 5, U32), ("srt_sum", 8, U32), ("art_sum", 9, U32),
                     ("srt_count", 12, U32), ("art_count", 13, U32),
                     ("counts_peer_tx", 14, MSG, "TcpPerfCountsPeer"), ("counts_peer_rx", 15, MSG, "TcpPerfCountsPeer"),
                     ("total_retrans_count", 16, U32), ("syn_count", 17, U32), ("synack_count", 18, U32),
                     ("cit_max", 19, U32), ("cit_sum", 20, U32), ("cit_count", 21, U32), ("fin_count", 22, U32)],
    "L7PerfStats": [("request_count", 1, U32), ("response_count", 2, U32), ("err_client_count", 3, U32),
                    ("err_server_count", 4, U32), ("err_timeout", 5, U32), ("rrt_count", 6, U32),
                    ("rrt_sum", 7, U64), ("rrt_max", 8, U32), ("tls_rtt", 9, U32)],
    "FlowPerfStats": [("tcp", 1, MSG, "TCPPerfStats"), ("l7", 2, MSG, "L7PerfStats"), ("l4_protocol", 3, U32),
                      ("l7_protocol", 4, U32), ("l7_failed_count", 5, U32)],
    "Flow": [("flow_key", 1, MSG, "FlowKey"), ("metrics_peer_src", 2, MSG, "FlowMetricsPeer"),
             ("metrics_peer_dst", 3, MSG, "FlowMetricsPeer"), ("Tunnel", 4, MSG, "TunnelField"),
             ("flow_id", 5, U64), ("start_time", 6, U64), ("end_time", 7, U64), ("duration", 8, U64),
             ("vlan", 10, U32), ("eth_type", 11, U32), ("has_perf_stats", 12, U32),
             ("perf_stats", 13, MSG, "FlowPerfStats"), ("close_type", 14, U32), ("signal_source", 15, U32),
             ("is_active_service", 16, U32), ("queue_hash", 17, U32), ("is_new_flow", 18, U32),
             ("tap_side", 19, U32), ("syn_seq", 20, U32), ("synack_seq", 21, U32),
             ("last_keepalive_seq", 22, U32), ("last_keepalive_ack", 23, U32), ("acl_gids", 24, U32, None, True),
             ("direction_score", 25, U32), ("request_domain", 26, STR),
             ("aggregated_flow_ids", 27, U64, None, True), ("init_ipid", 28, U32)],
    "TaggedFlow": [("flow", 1, MSG, "Flow")],
    "AppProtoHead": [("proto", 1, U32), ("msg_type", 2, U32), ("rrt", 5, U64)],
    "AppProtoLogsBaseInfo": [
        ("start_time", 1, U64), ("end_time", 2, U64), ("flow_id", 3, U64), ("tap_port", 4, U64),
        ("vtap_id", 5, U32), ("tap_type", 6, U32), ("is_ipv6", 7, U32), ("tap_side", 8, U32),
        ("head", 9, MSG, "AppProtoHead"), ("mac_src", 10, U64), ("mac_dst", 11, U64), ("ip_src", 12, U32),
        ("ip_dst", 13, U32), ("ip6_src", 14, BYT), ("ip6_dst", 15, BYT), ("l3_epc_id_src", 16, I32),
        ("l3_epc_id_dst", 17, I32), ("port_src", 18, U32), ("port_dst", 19, U32), ("protocol", 20, U32),
        ("is_vip_interface_src", 21, U32), ("is_vip_interface_dst", 22, U32), ("req_tcp_seq", 23, U32),
        ("resp_tcp_seq", 24, U32), ("process_id_0", 25, U32), ("process_id_1", 26, U32),
        ("process_kname_0", 27, STR), ("process_kname_1", 28, STR), ("syscall_trace_id_request", 29, U64),
        ("syscall_trace_id_response", 30, U64), ("syscall_trace_id_thread_0", 31, U32),
        ("syscall_trace_id_thread_1", 32, U32), ("syscall_cap_seq_0", 33, U32), ("syscall_cap_seq_1", 34, U32),
        ("gpid_0", 35, U32), ("gpid_1", 36, U32), ("syscall_coroutine_0", 39, U64),
        ("syscall_coroutine_1", 40, U64), ("pod_id_0", 41, U32), ("pod_id_1", 42, U32), ("biz_type", 43, U32)],
    "L7Request": [("req_type", 1, STR), ("domain", 2, STR), ("resource", 3, STR), ("endpoint", 4, STR)],
    "L7Response": [("status", 1, U32), ("code", 2, I32), ("exception", 3, STR), ("result", 4, STR)],
    "TraceInfo": [("trace_id", 1, STR), ("span_id", 2, STR), ("parent_span_id", 3, STR),
                  ("trace_ids", 4, STR, None, True)],
    "ExtendedInfo": [("service_name", 1, STR), ("client_ip", 2, STR), ("request_id", 3, U32),
                     ("x_request_id_0", 4, STR), ("http_user_agent", 6, STR), ("http_referer", 7, STR),
                     ("rpc_service", 8, STR), ("protocol_str", 9, STR), ("x_request_id_1", 10, STR),
                     ("attribute_names", 16, STR, None, True), ("attribute_values", 17, STR, None, True),
                     ("metrics_names", 18, STR, None, True), ("metrics_values", 19, DBL, None, True)],
    "AppProtoLogsData": [("base", 1, MSG, "AppProtoLogsBaseInfo"), ("req_len", 9, I32), ("resp_len", 10, I32),
                         ("req", 11, MSG, "L7Request"), ("resp", 12, MSG, "L7Response"), ("version", 13, STR),
                         ("trace_info", 14, MSG, "TraceInfo"), ("ext_info", 15, MSG, "ExtendedInfo"),
                         ("row_effect", 16, U32), ("direction_score", 17, U32), ("flags", 18, U32),
                         ("captured_request_byte", 19, U32), ("captured_response_byte", 20, U32),
                         ("biz_code", 21, STR), ("biz_scenario", 22, STR), ("biz_response_code", 23, STR)],
    # ---- message/metric.proto:14-195 ----
    "MiniField": [("ip", 1, BYT), ("ip1", 2, BYT), ("global_thread_id", 3, U32), ("is_ipv6", 4, U32),
                  ("l3_epc_id", 5, I32), ("l3_epc_id1", 6, I32), ("mac", 7, U64), ("mac1", 8, U64),
                  ("direction", 9, U32), ("tap_side", 10, U32), ("protocol", 11, U32), ("acl_gid", 12, U32),
                  ("server_port", 13, U32), ("vtap_id", 14, U32), ("tap_port", 15, U64), ("tap_type", 16, U32),
                  ("l7_protocol", 17, U32), ("gpid", 20, U32), ("gpid1", 21, U32), ("signal_source", 22, U32),
                  ("app_service", 23, STR), ("app_instance", 24, STR), ("endpoint", 25, STR), ("pod_id", 27, U32),
                  ("biz_type", 28, U32)],
    "MiniTag": [("field", 1, MSG, "MiniField"), ("code", 2, U64)],
    "Traffic": [("packet_tx", 1, U64), ("packet_rx", 2, U64), ("byte_tx", 3, U64), ("byte_rx", 4, U64),
                ("l3_byte_tx", 5, U64), ("l3_byte_rx", 6, U64), ("l4_byte_tx", 7, U64), ("l4_byte_rx", 8, U64),
                ("new_flow", 9, U64), ("closed_flow", 10, U64), ("l7_request", 11, U32), ("l7_response", 12, U32),
                ("syn", 13, U32), ("synack", 14, U32), ("direction_score", 15, U32)],
    "Latency": [("rtt_max", 1, U32), ("rtt_client_max", 2, U32), ("rtt_server_max", 3, U32), ("srt_max", 4, U32),
                ("art_max", 5, U32), ("rrt_max", 6, U32), ("cit_max", 19, U32), ("rtt_sum", 7, U64),
                ("rtt_client_sum", 8, U64), ("rtt_server_sum", 9, U64), ("srt_sum", 10, U64), ("art_sum", 11, U64),
                ("rrt_sum", 12, U64), ("cit_sum", 20, U64), ("rtt_count", 13, U32), ("rtt_client_count", 14, U32),
                ("rtt_server_count", 15, U32), ("srt_count", 16, U32), ("art_count", 17, U32),
                ("rrt_count", 18, U32), ("cit_count", 21, U32)],
    "Performance": [("retrans_tx", 1, U64), ("retrans_rx", 2, U64), ("zero_win_tx", 3, U64), ("zero_win_rx", 4, U64),
                    ("retrans_syn", 5, U32), ("retrans_synack", 6, U32)],
    "Anomaly": [("client_rst_flow", 1, U64), ("server_rst_flow", 2, U64), ("server_syn_miss", 3, U64),
                ("client_ack_miss", 4, U64), ("client_half_close_flow", 5, U64), ("server_half_close_flow", 6, U64),
                ("client_source_port_reuse", 7, U64), ("client_establish_reset", 8, U64), ("server_reset", 9, U64),
                ("server_queue_lack", 10, U64), ("server_establish_reset", 11, U64), ("tcp_timeout", 12, U64),
                ("l7_client_error", 13, U32), ("l7_server_error", 14, U32), ("l7_timeout", 15, U32),
                ("client_ooo", 20, U64), ("server_ooo", 21, U64)],
    "FlowLoad": [("load", 1, U64)],
    "FlowMeter": [("traffic", 1, MSG, "Traffic"), ("latency", 2, MSG, "Latency"),
                  ("performance", 3, MSG, "Performance"), ("anomaly", 4, MSG, "Anomaly"),
                  ("flow_load", 5, MSG, "FlowLoad")],
    "UsageMeter": [("packet_tx", 1, U64), ("packet_rx", 2, U64), ("byte_tx", 3, U64), ("byte_rx", 4, U64),
                   ("l3_byte_tx", 5, U64), ("l3_byte_rx", 6, U64), ("l4_byte_tx", 7, U64), ("l4_byte_rx", 8, U64)],
    "AppTraffic": [("request", 1, U32), ("response", 2, U32), ("direction_score", 3, U32)],
    "AppLatency": [("rrt_max", 1, U32), ("rrt_sum", 2, U64), ("rrt_count", 3, U32)],
    "AppAnomaly": [("client_error", 1, U32), ("server_error", 2, U32), ("timeout", 3, U32)],
    "AppMeter": [("traffic", 1, MSG, "AppTraffic"), ("latency", 2, MSG, "AppLatency"), ("anomaly", 3, MSG, "AppAnomaly")],
    "Meter": [("meter_id", 1, U32), ("flow", 2, MSG, "FlowMeter"), ("usage", 3, MSG, "UsageMeter"),
              ("app", 4, MSG, "AppMeter")],
    "Document": [("timestamp", 1, U32), ("tag", 2, MSG, "MiniTag"), ("meter", 3, MSG, "Meter"), ("flags", 4, U32)],
}

_CLASSES = None


def classes():
    global _CLASSES
    if _CLASSES is None:
        fd = descriptor_pb2.FileDescriptorProto()
        fd.name = "flow_log_test.proto"
        fd.package = "flow_log_test"
        fd.syntax = "proto3"
        for mname, fields in SCHEMA.items():
            m = fd.message_type.add()
            m.name = mname
            for f in fields:
                name, num, typ = f[0], f[1], f[2]
                sub = f[3] if len(f) > 3 else None
                rep = f[4] if len(f) > 4 else False
                fp = m.field.add()
                fp.name, fp.number, fp.type = name, num, typ
                fp.label = T.LABEL_REPEATED if rep else T.LABEL_OPTIONAL
                if typ == MSG:
                    fp.type_name = ".flow_log_test." + sub
        pool = descriptor_pool.DescriptorPool()
        pool.Add(fd)
        _CLASSES = {n: message_factory.GetMessageClass(pool.FindMessageTypeByName("flow_log_test." + n))
                    for n in SCHEMA}
    return _CLASSES


def frame(records):
    """[bytes] -> one frame payload in the agent's {u32 LE len, pb}* framing."""
    out = bytearray()
    for r in records:
        out += len(r).to_bytes(4, "little") + r
    return bytes(out)
