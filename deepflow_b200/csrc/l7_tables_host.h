/*
 * l7_tables_host.h — host-side construction of the constant strings the L7 path can emit (protocol names,
 * pseudo attribute names, exception descriptions) and the lookup tables over them.  Used by l7.cu (uploads
 * the blob) and by tests/harness (same tables on the host).
 */
#pragma once
#include <string>
#include <vector>

#include "l7_row.cuh"

namespace dfgpu {

// fills `cs` (string by id) and the id-valued members of `tb`; the caller lays `cs` out as blob + offsets
inline void l7_build_tables(std::vector<std::string>& cs, L7Tables& tb) {
  cs.assign(CS_FIXED_COUNT, std::string());
  cs[CS_ATTR_UA] = "http_user_agent";
  cs[CS_ATTR_REFERER] = "http_referer";
  cs[CS_ATTR_RPC] = "rpc_service";
  cs[CS_FT_APP_SERVICE] = "app_service";
  cs[CS_FT_ENDPOINT] = "endpoint";
  cs[CS_FT_APP_INSTANCE] = "app_instance";
  cs[CS_PROTO_NA] = "N/A";
  cs[CS_PROTO_TLS] = "TLS";
  cs[CS_DNS_AVAILABLE] = "available for assignment";
  cs[CS_DNS_PRIVATE] = "Private Use";
  cs[CS_KAFKA_UNKNOWN] = DFGPU_EXC_KAFKA_UNKNOWN_SERVER_ERROR;
  for (int i = 0; i < 256; i++) tb.proto_cs[i][0] = tb.proto_cs[i][1] = CS_PROTO_NA;
  struct PN { int id; const char* name; };
  static const PN protos[] = {  // L7Protocol.String, datatype/flow.go:623-803
      {20, "HTTP"}, {21, "HTTP2"}, {40, "Dubbo"}, {41, "gRPC"}, {43, "SofaRPC"}, {44, "FastCGI"}, {45, "bRPC"},
      {46, "Tars"}, {47, "SOME/IP"}, {48, "ISO-8583"}, {49, "TRIPLE"}, {50, "NetSign"}, {60, "MySQL"},
      {61, "PostgreSQL"}, {62, "Oracle"}, {63, "Dameng"}, {80, "Redis"}, {81, "MongoDB"}, {82, "Memcached"},
      {100, "Kafka"}, {101, "MQTT"}, {102, "AMQP"}, {103, "OpenWire"}, {104, "NATS"}, {105, "Pulsar"},
      {106, "ZMTP"}, {107, "RocketMQ"}, {108, "WebSphereMQ"}, {120, "DNS"}, {127, "Custom"}};
  for (auto& pn : protos) {
    tb.proto_cs[pn.id][0] = (uint16_t)cs.size();
    cs.push_back(pn.name);
    tb.proto_cs[pn.id][1] = (uint16_t)cs.size();
    cs.push_back(std::string(pn.name) + "_TLS");
  }
  tb.proto_cs[121][0] = tb.proto_cs[121][1] = CS_PROTO_TLS;
  auto add_table = [&](const char* const* t, int n) {
    uint32_t base = (uint32_t)cs.size();
    for (int i = 0; i < n; i++) cs.push_back(t[i]);
    return base;
  };
  tb.cs_http = add_table(dfgpu_exc_http, DFGPU_EXC_HTTP_LEN);
  tb.cs_dns = add_table(dfgpu_exc_dns, DFGPU_EXC_DNS_LEN);
  tb.cs_dubbo = add_table(dfgpu_exc_dubbo, DFGPU_EXC_DUBBO_LEN);
  tb.cs_kafka = add_table(dfgpu_exc_kafka, DFGPU_EXC_KAFKA_LEN);
  tb.cs_mqtt3 = add_table(dfgpu_exc_mqtt_v3, DFGPU_EXC_MQTT_V3_LEN);
  tb.cs_mqtt5 = add_table(dfgpu_exc_mqtt_v5, DFGPU_EXC_MQTT_V5_LEN);
}
inline void l7_layout_strings(const std::vector<std::string>& cs, std::vector<uint8_t>& blob, std::vector<uint32_t>& off) {
  blob.clear();
  off.clear();
  for (auto& s : cs) {
    off.push_back((uint32_t)blob.size());
    blob.insert(blob.end(), s.begin(), s.end());
  }
  off.push_back((uint32_t)blob.size());
}

}  // namespace dfgpu
