/*
 * kg_side.cuh — SmartEncoding for ONE endpoint of a record, register-only (device code of l7_kg.cu).
 * Same algorithm as kg.cuh's kg_fill (KnowledgeGraph.fill, flow_log/log_data/l4_flow_log.go:536-728;
 * ServiceTable.QueryPodService / QueryCustomService libs/grpc/grpc_service.go:138-323;
 * QueryIPV4Infos / QueryIPV6Infos grpc_platformdata.go:292-326), restated so that nothing is passed
 * by reference across a call and no byte array is indexed dynamically:
 *   - an IPv6 address is two u64 (bytes 0..7, 8..15, packed little-endian like t6_hash does);
 *   - a probe returns (found << 32 | value) in a register;
 *   - a lookup that can end in three kinds of answer (exact info, CIDR, epc base region) returns a
 *     64-bit "hit" = kind << 32 | index, applied to the output by the caller.
 * ncu, round 1: the by-reference InfoVal / IpAddr structs of kg.cuh cost the side-wise kernel a 448-byte
 * stack frame and 1.8 GB of local-memory traffic per 625 k records, half of its DRAM writes.
 * The first probe of the chains that do not depend on each other (gpid -> pod, pod -> info,
 * mac -> info, epc+ip -> info) is loaded before any of them is consumed: ~4 loads in flight per lane
 * instead of one.
 */
#pragma once
#include <cstddef>

#include "kg.cuh"

namespace dfgpu {
namespace kgs {

struct IpV {  // one endpoint address, by value
  uint32_t ip4;
  uint32_t len6;    // wire length of the v6 bytes (0 when absent)
  uint64_t lo, hi;  // v6 bytes 0..7 / 8..15, zero padded
};

// ---- Go net.IP predicates on the packed form (kg.cuh go_to4 / is_loopback / is_local_ip) ----------
__device__ __forceinline__ bool v4mapped(uint64_t lo, uint64_t hi) { return lo == 0 && (uint32_t)hi == 0xffff0000u; }
__device__ __forceinline__ bool to4(const IpV a, uint32_t& v4) {  // v4: byte i of the address in bits 8i..8i+7
  if (a.len6 == 4) { v4 = (uint32_t)a.lo; return true; }
  if (a.len6 == 16 && v4mapped(a.lo, a.hi)) { v4 = (uint32_t)(a.hi >> 32); return true; }
  return false;
}
__device__ __forceinline__ bool v4_global_unicast(uint32_t v) {  // bytes a,b,c,d = v, v>>8, v>>16, v>>24
  const uint32_t a = v & 0xff, b = (v >> 8) & 0xff;
  if (v == 0xffffffffu || v == 0) return false;  // bcast, unspecified
  if (a == 127) return false;                    // loopback
  if ((a & 0xf0) == 0xe0) return false;          // multicast
  if (a == 169 && b == 254) return false;        // link-local unicast
  return true;
}
__device__ __forceinline__ bool loopback(bool v6, const IpV a) {
  if (!v6) return (a.ip4 & 0xFF000000u) == 0x7F000000u;
  uint32_t v4;
  if (to4(a, v4)) return (v4 & 0xff) == 127;
  if (a.len6 != 16) return false;
  return a.lo == 0 && a.hi == 1ull << 56;
}
__device__ __forceinline__ bool local_ip(bool v6, const IpV a) {
  if (!v6) return !v4_global_unicast(__byte_perm(a.ip4, 0, 0x0123));  // ip4 is host order: a = top byte
  if (a.len6 != 4 && a.len6 != 16) return true;
  uint32_t v4;
  if (to4(a, v4)) return !v4_global_unicast(v4);
  if ((a.lo | a.hi) == 0) return true;
  if (a.lo == 0 && a.hi == 1ull << 56) return true;
  const uint32_t b0 = (uint32_t)a.lo & 0xff, b1 = ((uint32_t)a.lo >> 8) & 0xff;
  if (b0 == 0xff) return true;
  if (b0 == 0xfe && (b1 & 0xc0) == 0x80) return true;
  return false;
}

// ---- probes ------------------------------------------------------------------------------------------
constexpr uint64_t FOUND = 1ull << 32;

struct Pre {  // the first entry of a probe sequence, loaded ahead of its use
  uint4 e;
};
__device__ __forceinline__ Pre t64_pre(const DevDict& d, uint32_t tid, uint64_t key, bool want) {
  Pre r;
  const uint32_t s = t64_slot(tid, key, d.t64_mask);
  r.e = want ? __ldg(reinterpret_cast<const uint4*>(d.t64 + s)) : make_uint4(0, 0, 0, TID_EMPTY);
  return r;
}
// the rest of the sequence (linear probing past a foreign key): rare, out of line
static __device__ __noinline__ uint64_t t64_more(const T64Entry* t, uint32_t mask, uint32_t tid, uint64_t key) {
  uint32_t s = t64_slot(tid, key, mask);
  for (;;) {
    s = (s + 1) & mask;
    const uint4 raw = __ldg(reinterpret_cast<const uint4*>(t + s));
    if (raw.w == TID_EMPTY) return 0;
    if (raw.w == tid && ((uint64_t)raw.y << 32 | raw.x) == key) return FOUND | raw.z;
  }
}
__device__ __forceinline__ uint64_t t64_fin(const DevDict& d, uint32_t tid, uint64_t key, Pre pre) {
  const uint4 raw = pre.e;
  if (raw.w == TID_EMPTY) return 0;
  if (raw.w == tid && ((uint64_t)raw.y << 32 | raw.x) == key) return FOUND | raw.z;
  return t64_more(d.t64, d.t64_mask, tid, key);
}
__device__ __forceinline__ uint64_t t64_get(const DevDict& d, uint32_t tid, uint64_t key) {
  return t64_fin(d, tid, key, t64_pre(d, tid, key, true));
}
__device__ __forceinline__ uint32_t t64_z(const DevDict& d, uint32_t tid, uint64_t key) {
  return (uint32_t)t64_get(d, tid, key);  // value, or 0 when absent (FOUND is bit 32)
}
static_assert(sizeof(T6Entry) == 32 && offsetof(T6Entry, epc) == 16 && offsetof(T6Entry, port) == 20 &&
                  offsetof(T6Entry, proto) == 22 && offsetof(T6Entry, tid) == 23 && offsetof(T6Entry, val) == 24,
              "T6Entry is read as two uint4");
static __device__ __noinline__ uint64_t t6_get(const T6Entry* t, uint32_t mask, uint32_t tid, uint64_t lo, uint64_t hi,
                                        int32_t epc, uint32_t port, uint32_t proto) {
  const uint64_t c = (uint64_t)(uint32_t)epc << 32 | (uint64_t)port << 16 | (uint64_t)proto << 8 | tid;
  uint32_t s = (uint32_t)mix64(lo ^ mix64(hi ^ mix64(c))) & mask;  // = t6_hash
  const uint32_t meta = port | proto << 16 | tid << 24;
  for (;;) {
    const uint4* q = reinterpret_cast<const uint4*>(t + s);
    const uint4 m = __ldg(q + 1);
    if ((m.y >> 24) == 0) return 0;
    if (m.y == meta && (int32_t)m.x == epc) {
      const uint4 a = __ldg(q);
      if (((uint64_t)a.y << 32 | a.x) == lo && ((uint64_t)a.w << 32 | a.z) == hi) return FOUND | m.z;
    }
    s = (s + 1) & mask;
  }
}
__device__ __forceinline__ uint32_t t6_z(const DevDict& d, uint32_t tid, const IpV a, int32_t epc, uint32_t port,
                                         uint32_t proto) {
  return (uint32_t)t6_get(d.t6, d.t6_mask, tid, a.lo, a.hi, epc, port, proto);
}

// ---- "hit": where the endpoint's Info comes from ------------------------------------------------------
enum : uint32_t { H_NONE = 0, H_INFO = 1, H_CIDR4 = 2, H_CIDR6 = 3, H_REGION = 4 };
__device__ __forceinline__ uint64_t hit(uint32_t kind, uint32_t v) { return (uint64_t)kind << 32 | v; }
__device__ __forceinline__ uint32_t hit_kind(uint64_t h) { return (uint32_t)(h >> 32); }

__device__ __forceinline__ uint64_t epc_base(const DevDict& d, int32_t epc) {  // queryEpcIDBaseInfo :686-700
  if (epc == EPC_UNKNOWN) return hit(H_REGION, d.region_id);
  const uint64_t r = t64_get(d, TID_EPC_BASE, (uint64_t)(uint32_t)epc);
  return r ? hit(H_REGION, (uint32_t)r) : 0;
}
// QueryIPV4Infos :292-308 after the exact (epc, ip) probe missed: first containing CIDR, then epc base
static __device__ __noinline__ uint64_t ip4_rest(const DevDict& d, int32_t epc, uint32_t ip) {
  const uint64_t r = t64_get(d, TID_CIDR4, (uint64_t)(uint32_t)epc);
  if (r) {
    const DRange rg = d.cidr4_ranges[(uint32_t)r];
    for (uint32_t i = 0; i < rg.count; i++) {
      const DCidr4& c = d.cidr4[rg.start + i];
      if ((ip & c.mask) == c.net) return hit(H_CIDR4, rg.start + i);
    }
  }
  return epc_base(d, epc);
}
__device__ __forceinline__ uint64_t ip4_key(int32_t epc, uint32_t ip) { return (uint64_t)(int64_t)epc << 32 | (uint64_t)ip; }
__device__ __forceinline__ uint64_t query_ip4(const DevDict& d, int32_t epc, uint32_t ip, Pre pre) {
  if (epc == EPC_FROM_INTERNET) return 0;
  const uint64_t r = t64_fin(d, TID_IP4, ip4_key(epc, ip), pre);
  if (r) return hit(H_INFO, (uint32_t)r);
  return ip4_rest(d, epc, ip);
}
// bytes 0..n-1 of a packed u64 all ones (n <= 8)
__device__ __forceinline__ uint64_t byte_mask(uint32_t bits) {  // prefix mask of `bits` (< 64) leading address bits
  const uint32_t full = bits >> 3, rem = bits & 7;
  uint64_t m = full ? ~0ull >> (64 - 8 * full) : 0ull;
  if (rem) m |= (uint64_t)(uint8_t)(0xff << (8 - rem)) << (8 * full);
  return m;
}
// QueryIPV6Infos :310-326
static __device__ __noinline__ uint64_t query_ip6(const DevDict& d, int32_t epc, uint64_t lo, uint64_t hi, uint32_t len6) {
  if (epc == EPC_FROM_INTERNET) return 0;
  const uint64_t r = t6_get(d.t6, d.t6_mask, T6_IP6, lo, hi, epc, 0, 0);
  if (r) return hit(H_INFO, (uint32_t)r);
  // net.IPNet.Contains: only a 16-byte, non-v4-mapped address can match a v6 net
  if (len6 == 16 && !v4mapped(lo, hi)) {
    const uint64_t rr = t64_get(d, TID_CIDR6, (uint64_t)(uint32_t)epc);
    if (rr) {
      const DRange rg = d.cidr6_ranges[(uint32_t)rr];
      for (uint32_t i = 0; i < rg.count; i++) {
        const DCidr6& c = d.cidr6[rg.start + i];
        uint64_t nlo = 0, nhi = 0;
        for (int b = 0; b < 8; b++) {
          nlo |= (uint64_t)c.net[b] << (8 * b);
          nhi |= (uint64_t)c.net[8 + b] << (8 * b);
        }
        const uint32_t pl = c.plen > 128 ? 128 : c.plen;
        const uint64_t mlo = pl >= 64 ? ~0ull : byte_mask(pl);
        const uint64_t mhi = pl >= 128 ? ~0ull : pl > 64 ? byte_mask(pl - 64) : 0ull;
        if (((lo ^ nlo) & mlo) == 0 && ((hi ^ nhi) & mhi) == 0) return hit(H_CIDR6, rg.start + i);
      }
    }
  }
  return epc_base(d, epc);
}
__device__ __forceinline__ uint64_t reget(const DevDict& d, bool v6, const IpV a, int32_t epc, Pre pre4) {
  return v6 ? query_ip6(d, epc, a.lo, a.hi, a.len6) : query_ip4(d, epc, a.ip4, pre4);
}

static_assert(sizeof(DInfo) == 64 && offsetof(DInfo, l2_epc) == 4 && offsetof(DInfo, host_id) == 8 &&
                  offsetof(DInfo, region_id) == 12 && offsetof(DInfo, device_type) == 16 &&
                  offsetof(DInfo, device_id) == 20 && offsetof(DInfo, subnet_id) == 24 &&
                  offsetof(DInfo, pod_node_id) == 28 && offsetof(DInfo, pod_ns_id) == 32 &&
                  offsetof(DInfo, pod_group_id) == 36 && offsetof(DInfo, pod_id) == 40 &&
                  offsetof(DInfo, pod_cluster_id) == 44 && offsetof(DInfo, az_id) == 48 &&
                  offsetof(DInfo, pod_group_type) == 52 && offsetof(DInfo, dup) == 53,
              "DInfo is read as four uint4");
// the pod id an answer carries (QueryPodService :150-158 compares it)
__device__ __forceinline__ uint32_t hit_pod_id(const DevDict& d, uint64_t h) {
  return hit_kind(h) == H_INFO ? d.infos[(uint32_t)h].pod_id : 0u;
}
// queryMacInfo :613-623 (DupCount > 0 => nil): info index, from an already issued first probe
__device__ __forceinline__ uint64_t query_mac(const DevDict& d, uint64_t l3mac, Pre pre) {
  const uint64_t r = t64_fin(d, TID_MAC, l3mac, pre);
  if (!r) return 0;
  if (d.infos[(uint32_t)r].dup) return 0;
  return hit(H_INFO, (uint32_t)r);
}

// :677-704 copy of *grpc.Info (or of the CIDR / epc-base answer) into the row
__device__ __forceinline__ void apply(const DevDict& d, uint64_t h, KGSide& o) {
  const uint32_t kind = hit_kind(h), v = (uint32_t)h;
  if (kind == H_INFO) {
    const uint4* q = reinterpret_cast<const uint4*>(d.infos + v);
    const uint4 a = __ldg(q), b = __ldg(q + 1), c = __ldg(q + 2), e = __ldg(q + 3);
    o.host_id = (uint16_t)a.z;
    o.region_id = (uint16_t)a.w;
    o.l3_device_type = (uint8_t)b.x;
    o.l3_device_id = b.y;
    o.subnet_id = (uint16_t)b.z;
    o.pod_node_id = b.w;
    o.pod_ns_id = (uint16_t)c.x;
    o.pod_group_id = c.y;
    o.pod_id = c.z;
    o.pod_cluster_id = (uint16_t)c.w;
    o.az_id = (uint16_t)e.x;
    o.pod_group_type = (uint8_t)e.y;
  } else if (kind == H_CIDR4) {
    const DCidr4& c = d.cidr4[v];
    o.subnet_id = (uint16_t)c.subnet_id;
    o.region_id = (uint16_t)c.region_id;
    o.az_id = (uint16_t)c.az_id;
  } else if (kind == H_CIDR6) {
    const DCidr6& c = d.cidr6[v];
    o.subnet_id = (uint16_t)c.subnet_id;
    o.region_id = (uint16_t)c.region_id;
    o.az_id = (uint16_t)c.az_id;
  } else if (kind == H_REGION) {
    o.region_id = (uint16_t)v;
  }
}

// ---- services (:713-727) ------------------------------------------------------------------------------
// QueryPodService (grpc_service.go:138-193) and QueryCustomService (:195-323) are priority lists of
// look-ups whose keys all follow from the endpoint's Info: every first probe is issued before any is
// consumed (up to 11 sectors in flight per lane, executed by the whole warp), then the answers are taken
// in the reference's order.  Only customServicePodService (keyed by the pod-service answer) and the
// "does the address still belong to this pod" re-check (:150-158) stay dependent.
static __device__ __noinline__ uint32_t custom_service_ip6(const DevDict& d, int32_t epc, uint64_t lo, uint64_t hi,
                                                    uint32_t server_port, uint32_t sp) {
  uint32_t sid = 0;
  if (d.counts6[T6_CS_IP6PORT]) {
    sid = (uint32_t)t6_get(d.t6, d.t6_mask, T6_CS_IP6PORT, lo, hi, epc, server_port, sp);
    if (sid) return sid;
  }
  if (d.counts6[T6_CS_IP6] && server_port != 0) {
    sid = (uint32_t)t6_get(d.t6, d.t6_mask, T6_CS_IP6, lo, hi, epc, 0, 0);
    if (sid) return sid;
  }
  if (d.counts6[T6_CS_IP6PORT]) {
    sid = (uint32_t)t6_get(d.t6, d.t6_mask, T6_CS_IP6PORT, lo, hi, 0, server_port, sp);
    if (sid) return sid;
  }
  if (d.counts6[T6_CS_IP6] && server_port != 0) sid = (uint32_t)t6_get(d.t6, d.t6_mask, T6_CS_IP6, lo, hi, 0, 0, 0);
  return sid;
}

struct SideIn;
__device__ __forceinline__ void services(const DevDict& d, const SideIn& in, KGSide& o, Pre pre_ip4);

struct SideIn {
  bool v6, vip;
  int32_t l3_epc;
  IpV ip;
  uint64_t mac;
  uint32_t gpid;
  uint32_t pod;        // after kg_peer_pods
  uint8_t tag_source;  // TS_PEER when the pod was borrowed
  uint16_t vtap_id;
  uint16_t port;       // side 1: server port; side 0: 0 (:714-727)
  uint32_t tap_side;
  uint8_t protocol;
  bool server_side;    // side 1 also matches pod-node services (:717-719)
};

__device__ __forceinline__ void fill_side(const DevDict& d, const SideIn& in, KGSide& o) {
  o.region_id = o.az_id = o.host_id = 0;
  o.l3_device_type = 0;
  o.l3_device_id = o.pod_node_id = 0;
  o.pod_ns_id = 0;
  o.pod_group_id = 0;
  o.pod_group_type = 0;
  o.pod_id = 0;
  o.pod_cluster_id = 0;
  o.epc_id = 0;
  o.subnet_id = 0;
  o.service_id = 0;
  o.tag_source = in.tag_source;
  uint32_t pod = in.pod;
  bool by_mac = in.vip, by_agent = false;
  if (in.tap_side == 4) {  // Local :570-585
    if (local_ip(in.v6, in.ip)) {
      if (in.mac != 0) by_mac = true; else by_agent = true;
    }
  } else if (in.tap_side == 41 || in.tap_side == 42) {  // ClientProcess / ServerProcess :586-598
    if (in.mac != 0) by_mac = true;
    else if (local_ip(in.v6, in.ip)) by_agent = true;
  }
  const uint64_t l3mac = in.mac | (uint64_t)(int64_t)in.l3_epc << 48;
  // first probes of the independent chains, issued together
  const bool want_gp = in.gpid != 0 && pod == 0;
  const Pre pre_gp = t64_pre(d, TID_GPROC, in.gpid, want_gp);
  const Pre pre_pod = t64_pre(d, TID_POD, pod, pod != 0);
  const Pre pre_mac = t64_pre(d, TID_MAC, l3mac, by_mac || in.l3_epc > 0);
  const Pre pre_ip4 = t64_pre(d, TID_IP4, ip4_key(in.l3_epc, in.ip.ip4), !in.v6 && in.l3_epc != EPC_FROM_INTERNET);

  uint64_t h = 0;
  if (want_gp) {  // :601-614
    const uint64_t gi = t64_fin(d, TID_GPROC, in.gpid, pre_gp);
    if (gi) {
      const uint64_t vp = d.gp_vals[(uint32_t)gi];
      const uint32_t p = (uint32_t)vp;
      if (p != 0 && (uint16_t)(vp >> 32) == in.vtap_id) {
        pod = p;
        o.tag_source |= TS_GPID;
      }
    }
    if (pod != 0) {  // :617-624, the pod came from the gpid
      o.tag_source |= TS_PODID;
      const uint64_t r = t64_get(d, TID_POD, pod);
      if (r) h = hit(H_INFO, (uint32_t)r);
    }
  } else if (pod != 0) {  // :617-624
    o.tag_source |= TS_PODID;
    const uint64_t r = t64_fin(d, TID_POD, pod, pre_pod);
    if (r) h = hit(H_INFO, (uint32_t)r);
  }
  // one mac answer serves both :626-634 and the L2 epc :668-675
  uint64_t hmac = 0;
  bool mac_done = false;
  if (!h) {  // :626-666
    if (by_mac) {
      o.tag_source |= TS_MAC;
      hmac = query_mac(d, l3mac, pre_mac);
      mac_done = true;
      h = hmac;
    }
    if (!h && by_agent) {
      o.tag_source |= TS_AGENT;
      const uint64_t vi = t64_get(d, TID_VTAP, in.vtap_id);
      if (vi) {
        const DVtap& v = d.vtaps[(uint32_t)vi];
        if (v.is_v4) {
          const int32_t ve = v.epc;
          const uint32_t vip = v.ip4;
          h = query_ip4(d, ve, vip, t64_pre(d, TID_IP4, ip4_key(ve, vip), ve != EPC_FROM_INTERNET));
        } else {
          uint64_t lo = 0, hi = 0;
          for (int i = 0; i < 8; i++) {
            lo |= (uint64_t)v.ip6[i] << (8 * i);
            hi |= (uint64_t)v.ip6[8 + i] << (8 * i);
          }
          h = query_ip6(d, v.epc, lo, hi, v.has_ip6 ? 16u : 0u);
        }
      }
    }
    if (!h) {
      o.tag_source |= TS_EPCIP;
      h = reget(d, in.v6, in.ip, in.l3_epc, pre_ip4);
    }
  }
  if (in.l3_epc > 0) {  // :668-675 L2 epc through the mac table
    if (!mac_done) hmac = query_mac(d, l3mac, pre_mac);
    if (hmac) o.epc_id = d.infos[(uint32_t)hmac].l2_epc;
  }
  apply(d, h, o);  // :677-704
  o.l3_epc_id = in.l3_epc;
  services(d, in, o, pre_ip4);  // :713-727
}

__device__ __forceinline__ void services(const DevDict& d, const SideIn& in, KGSide& o, Pre pre_ip4) {
  const int32_t epc = in.l3_epc;
  const uint16_t port = in.port;
  const uint32_t sp = to_service_protocol(port == 0 ? (uint8_t)0 : in.protocol);  // both queries: port 0 => protocol 0
  const uint32_t ip4 = in.ip.ip4;
  // ---- keys + first probes -----------------------------------------------------------------------------
  const bool want_ps = o.l3_device_type == 11 || o.pod_id != 0 || (in.server_side && o.pod_node_id != 0);
  const bool ps_group = o.pod_id != 0, ps_cluster = o.pod_id == 0 && o.pod_node_id != 0 && port != 0;
  const uint32_t ps1_tid = ps_group ? TID_PS_GROUPPORT : TID_PS_CLUSTERPORT;
  const uint64_t ps1_key = key_podx_port(ps_group ? o.pod_group_id : o.pod_cluster_id, sp, port);
  const uint64_t ps2_key = key_epc_ip4_port(epc, ip4, port);
  const Pre p_ps1 = t64_pre(d, ps1_tid, ps1_key, want_ps && (ps_group || ps_cluster));
  const Pre p_ps2 = t64_pre(d, TID_PS_IP4PORT0 + sp, ps2_key, want_ps && !in.v6 && epc > 0);
  const bool cs = epc >= 0;
  const bool w_pg = cs && o.pod_group_id != 0 && d.counts[TID_CS_PODGROUP] != 0;
  const bool w_pod = cs && o.pod_id != 0 && d.counts[TID_CS_POD] != 0;
  const bool w_clp = cs && o.pod_cluster_id != 0 && port != 0 && d.counts[TID_CS_CLUSTERPORT] != 0;
  const bool w_cl = cs && o.pod_cluster_id != 0 && port != 0 && d.counts[TID_CS_CLUSTER] != 0;
  const uint32_t host_tid = o.l3_device_type == 1 ? TID_CS_CHOST : TID_CS_HOST;  // VMDevice / HostDevice
  const bool w_host = cs && o.l3_device_id != 0 &&
                      ((o.l3_device_type == 1 && d.counts[TID_CS_CHOST] != 0) ||
                       (o.l3_device_type == 6 && d.counts[TID_CS_HOST] != 0));
  const bool w_ipp = cs && !in.v6 && d.counts[TID_CS_IP4PORT0 + sp] != 0;
  const bool w_ip = cs && !in.v6 && d.counts[TID_CS_IP4] != 0 && port != 0;
  const uint64_t k_clp = key_podx_port(o.pod_cluster_id, sp, port), k_cl = (uint64_t)o.pod_cluster_id << 32;
  const uint64_t k_i1 = ps2_key, k_i2 = key_epc_ip4_port(epc, ip4, 0);  // genEpcIDIP4Key(.., protocol 0)
  const uint64_t k_i3 = key_epc_ip4_port(0, ip4, port), k_i4 = key_epc_ip4_port(0, ip4, 0);
  const Pre c_pg = t64_pre(d, TID_CS_PODGROUP, o.pod_group_id, w_pg);
  const Pre c_pod = t64_pre(d, TID_CS_POD, o.pod_id, w_pod);
  const Pre c_clp = t64_pre(d, TID_CS_CLUSTERPORT, k_clp, w_clp);
  const Pre c_cl = t64_pre(d, TID_CS_CLUSTER, k_cl, w_cl);
  const Pre c_host = t64_pre(d, host_tid, o.l3_device_id, w_host);
  const Pre c_i1 = t64_pre(d, TID_CS_IP4PORT0 + sp, k_i1, w_ipp);
  const Pre c_i2 = t64_pre(d, TID_CS_IP4, k_i2, w_ip);
  const Pre c_i3 = t64_pre(d, TID_CS_IP4PORT0 + sp, k_i3, w_ipp);
  const Pre c_i4 = t64_pre(d, TID_CS_IP4, k_i4, w_ip);

  // ---- QueryPodService :138-193 ---------------------------------------------------------------------
  uint32_t sid = 0;
  if (want_ps) {
    bool dead = false;
    if (ps_group) {
      sid = (uint32_t)t64_fin(d, ps1_tid, ps1_key, p_ps1);
      if (sid == 0) {  // :150-158 the address now belongs to another pod: no service
        const uint64_t h = reget(d, in.v6, in.ip, epc, pre_ip4);
        dead = h != 0 && hit_pod_id(d, h) != o.pod_id;
      }
    } else if (ps_cluster) {
      sid = (uint32_t)t64_fin(d, ps1_tid, ps1_key, p_ps1);
    }
    if (sid == 0 && !dead && epc > 0)
      sid = in.v6 ? t6_z(d, T6_PS_IP6PORT, in.ip, epc, port, sp)
                  : (uint32_t)t64_fin(d, TID_PS_IP4PORT0 + sp, ps2_key, p_ps2);
    if (dead) sid = 0;
  }
  o.service_id = sid;
  get_auto_instance(o.pod_id, in.gpid, o.pod_node_id, o.l3_device_id, o.subnet_id, o.l3_device_type, o.l3_epc_id,
                    o.auto_instance_id, o.auto_instance_type);
  // ---- QueryCustomService :195-323 ------------------------------------------------------------------
  uint32_t custom = 0;
  if (cs) {
    if (sid != 0 && d.counts[TID_CS_PODSVC]) custom = t64_z(d, TID_CS_PODSVC, sid);
    if (!custom && w_pg) custom = (uint32_t)t64_fin(d, TID_CS_PODGROUP, o.pod_group_id, c_pg);
    if (!custom && w_pod) custom = (uint32_t)t64_fin(d, TID_CS_POD, o.pod_id, c_pod);
    if (!custom && w_clp) custom = (uint32_t)t64_fin(d, TID_CS_CLUSTERPORT, k_clp, c_clp);
    if (!custom && w_cl) custom = (uint32_t)t64_fin(d, TID_CS_CLUSTER, k_cl, c_cl);
    if (!custom && w_host) custom = (uint32_t)t64_fin(d, host_tid, o.l3_device_id, c_host);
    if (!custom) {
      if (in.v6) {
        if (d.counts6[T6_CS_IP6PORT] | d.counts6[T6_CS_IP6])
          custom = custom_service_ip6(d, epc, in.ip.lo, in.ip.hi, port, sp);
      } else {
        if (w_ipp) custom = (uint32_t)t64_fin(d, TID_CS_IP4PORT0 + sp, k_i1, c_i1);
        if (!custom && w_ip) custom = (uint32_t)t64_fin(d, TID_CS_IP4, k_i2, c_i2);
        if (!custom && w_ipp) custom = (uint32_t)t64_fin(d, TID_CS_IP4PORT0 + sp, k_i3, c_i3);
        if (!custom && w_ip) custom = (uint32_t)t64_fin(d, TID_CS_IP4, k_i4, c_i4);
      }
    }
  }
  get_auto_service(custom, o.service_id, o.pod_group_id, in.gpid, o.pod_cluster_id, o.l3_device_id, o.subnet_id,
                   o.l3_device_type, o.pod_group_type, o.l3_epc_id, o.auto_service_id, o.auto_service_type);
}

}  // namespace kgs
}  // namespace dfgpu
