/*
 * tables.cuh — the SmartEncoding dictionary as GPU-resident, read-only hash
 * tables (one immutable snapshot per org, swapped by pointer between batches —
 * the same life-cycle as the reference's ReloadSlave pointer swap,
 * server/libs/grpc/grpc_platformdata.go:1166-1228).
 *
 * Layout (all in HBM, built once per dictionary version by dict_build.cu):
 *   T64   open-addressing table of 16-byte entries {key u64, val u32, tid u32};
 *         ALL of the reference's uint64/uint32-keyed Go maps live in this one
 *         table, told apart by `tid` (grpc_platformdata.go:147-209 and
 *         grpc_service.go:32-49 list the maps).  One probe = one 16-byte load
 *         = one 32-byte HBM sector.
 *   T6    same idea for the 20/24-byte IPv6-keyed maps, 32-byte entries.
 *   infos pool of 64-byte DInfo records (compact grpc.Info, :64-90).
 *   cidr4/cidr6 per-epc CIDR lists, longest prefix first (:1476-1491).
 * The probe helpers are __host__ __device__ so the host-side builder resolves
 * pod_ips (updatePodIps :1907-1936) through exactly the code the kernels run.
 */
#pragma once
#include <cstdint>

#ifdef __CUDACC__
#define DF_HD __host__ __device__ __forceinline__
#define DF_HD_NOINLINE __host__ __device__ __noinline__
#else
#define DF_HD inline
#define DF_HD_NOINLINE inline
#endif
// The dictionary probes and the record-wise SmartEncoding routine (kg.cuh kg_fill: the L4 pass and the
// flow_metrics pass) are compiled OUT OF LINE: inside the big row-building kernels they must not bloat the
// instruction stream (ncu: 31 % instruction-fetch stalls when everything was inlined).  DF_KG_INLINE (all
// inline, sides unrolled) is kept as a build option for experiments; the L7 path does not use these functions
// any more — its dedicated kernel (l7_kg.cu) runs the register-only, one-side-per-lane restatement in
// kg_side.cuh.
#ifdef DF_KG_INLINE
#define DF_KG_FN static DF_HD
#define DF_KG_SIDE_LOOP _Pragma("unroll")
#else
#define DF_KG_FN static DF_HD_NOINLINE
#define DF_KG_SIDE_LOOP _Pragma("unroll 1")
#endif

namespace dfgpu {

enum Tid : uint32_t {
  TID_IP4 = 0,        // epcIDIPV4Infos        key epc<<32|ip            -> info idx
  TID_MAC = 1,        // macInfos              key mac|epc<<48           -> info idx
  TID_POD = 2,        // podIDInfos            key pod id                -> info idx
  TID_GPROC = 3,      // gprocessInfos         key gpid                  -> gp_vals idx
  TID_VTAP_PID = 4,   // vtapIDProcessInfos    key vtap<<32|pid          -> gpid
  TID_VTAP = 5,       // vtapIdInfos           key vtap id               -> vtaps idx
  TID_EPC_BASE = 6,   // epcIDBaseInfos        key (u32)epc              -> region id
  TID_CIDR4 = 7,      // epcIDIPV4CidrInfos    key (u32)epc              -> cidr4 range idx
  TID_CIDR6 = 8,      // epcIDIPV6CidrInfos
  TID_PS_IP4PORT0 = 9,    // podServiceClusterIP4PortTable[ANY|TCP|UDP] (+protocol)
  TID_PS_CLUSTERPORT = 12,  // podServiceClusterIDPortTable
  TID_PS_GROUPPORT = 13,    // podServiceGroupIDPortTable
  TID_CS_IP4PORT0 = 14,     // customServiceClusterIP4PortTable[3]
  TID_CS_IP4 = 17,          // customServiceClusterIP4Table
  TID_CS_CLUSTERPORT = 18,  // customServiceClusterIDPortTable
  TID_CS_CLUSTER = 19,      // customServiceClusterIDTable
  TID_CS_PODSVC = 20,       // customServicePodService
  TID_CS_PODGROUP = 21,
  TID_CS_POD = 22,
  TID_CS_CHOST = 23,
  TID_CS_HOST = 24,
  TID_COUNT = 25,
  TID_EMPTY = 0xffffffffu
};
enum Tid6 : uint8_t {
  T6_IP6 = 1,          // epcIDIPV6Infos                   -> info idx
  T6_PS_IP6PORT = 2,   // podServiceClusterIP6PortTable    -> service id
  T6_CS_IP6PORT = 3,   // customServiceClusterIP6PortTable
  T6_CS_IP6 = 4,       // customServiceClusterIP6Table
  T6_COUNT = 5
};

struct T64Entry {
  uint64_t key;
  uint32_t val;
  uint32_t tid;
};
static_assert(sizeof(T64Entry) == 16, "T64Entry");

struct T6Entry {
  uint8_t ip6[16];
  int32_t epc;
  uint16_t port;
  uint8_t proto;
  uint8_t tid;  // 0 = empty
  uint32_t val;
  uint32_t pad;
};
static_assert(sizeof(T6Entry) == 32, "T6Entry");

// Compact grpc.Info (grpc_platformdata.go:64-90): only what KnowledgeGraph.fill,
// QueryPodService and the metrics path read.
struct DInfo {
  int32_t epc;
  int32_t l2_epc;
  uint32_t host_id;
  uint32_t region_id;
  uint32_t device_type;
  uint32_t device_id;
  uint32_t subnet_id;
  uint32_t pod_node_id;
  uint32_t pod_ns_id;
  uint32_t pod_group_id;
  uint32_t pod_id;
  uint32_t pod_cluster_id;
  uint32_t az_id;
  uint8_t pod_group_type;
  uint8_t dup;  // DupCount > 0
  uint8_t pad[2];
  uint32_t pad2[2];
};
static_assert(sizeof(DInfo) == 64, "DInfo");

struct DVtap {  // grpc.VtapInfo :115-125
  int32_t epc;
  uint32_t is_v4;
  uint32_t ip4;
  uint32_t has_ip6;
  uint8_t ip6[16];
};
struct DCidr4 {
  uint32_t net, mask;
  uint32_t subnet_id, region_id, az_id;
};
struct DCidr6 {
  uint8_t net[16];
  uint32_t plen;
  uint32_t subnet_id, region_id, az_id;
};
struct DRange {
  uint32_t start, count;
};

struct DevDict {
  const T64Entry* t64;
  uint32_t t64_mask;
  const T6Entry* t6;
  uint32_t t6_mask;
  const DInfo* infos;
  const uint64_t* gp_vals;
  const DVtap* vtaps;
  const DCidr4* cidr4;
  const DRange* cidr4_ranges;
  const DCidr6* cidr6;
  const DRange* cidr6_ranges;
  uint32_t counts[TID_COUNT];  // len(map) per tid, for the `len(table) > 0` guards
  uint32_t counts6[T6_COUNT];
  uint32_t region_id;    // t.regionID[org]
  uint32_t analyzer_id;  // t.analyzerID
};

DF_HD uint64_t mix64(uint64_t x) {
  x ^= x >> 33;
  x *= 0xff51afd7ed558ccdull;
  x ^= x >> 33;
  x *= 0xc4ceb9fe1a85ec53ull;
  x ^= x >> 33;
  return x;
}
DF_HD uint32_t t64_slot(uint32_t tid, uint64_t key, uint32_t mask) {
  return (uint32_t)mix64(key ^ ((uint64_t)tid * 0x9E3779B97F4A7C15ull)) & mask;
}

DF_KG_FN bool t64_find(const T64Entry* t, uint32_t mask, uint32_t tid, uint64_t key, uint32_t& val) {
  uint32_t s = t64_slot(tid, key, mask);
  for (;;) {
#if defined(__CUDA_ARCH__)
    const uint4 raw = __ldg(reinterpret_cast<const uint4*>(t + s));
    uint64_t k = (uint64_t)raw.y << 32 | raw.x;
    uint32_t v = raw.z, id = raw.w;
#else
    uint64_t k = t[s].key;
    uint32_t v = t[s].val, id = t[s].tid;
#endif
    if (id == TID_EMPTY) return false;
    if (id == tid && k == key) {
      val = v;
      return true;
    }
    s = (s + 1) & mask;
  }
}

DF_HD uint64_t t6_hash(const uint8_t* ip6, int32_t epc, uint16_t port, uint8_t proto, uint8_t tid) {
  uint64_t a = 0, b = 0;
  for (int i = 0; i < 8; i++) {
    a |= (uint64_t)ip6[i] << (8 * i);
    b |= (uint64_t)ip6[8 + i] << (8 * i);
  }
  uint64_t c = (uint64_t)(uint32_t)epc << 32 | (uint64_t)port << 16 | (uint64_t)proto << 8 | tid;
  return mix64(a ^ mix64(b ^ mix64(c)));
}
DF_KG_FN bool t6_find(const T6Entry* t, uint32_t mask, uint8_t tid, const uint8_t* ip6, int32_t epc,
                   uint16_t port, uint8_t proto, uint32_t& val) {
  uint32_t s = (uint32_t)t6_hash(ip6, epc, port, proto, tid) & mask;
  for (;;) {
    const T6Entry& e = t[s];
    if (e.tid == 0) return false;
    if (e.tid == tid && e.epc == epc && e.port == port && e.proto == proto) {
      bool same = true;
      for (int i = 0; i < 16; i++) same &= e.ip6[i] == ip6[i];
      if (same) {
        val = e.val;
        return true;
      }
    }
    s = (s + 1) & mask;
  }
}

// ---- lookups restating grpc_platformdata.go queries ------------------------------
constexpr int32_t EPC_FROM_INTERNET = -2;  // libs/datatype/endpoint.go:29
constexpr int32_t EPC_UNKNOWN = 0;         // :30

struct InfoVal {  // what KnowledgeGraph.fill copies out of *grpc.Info
  uint32_t region_id, az_id, host_id, device_type, device_id, pod_node_id, pod_ns_id, pod_group_id,
      pod_id, pod_cluster_id, subnet_id;
  int32_t l2_epc;
  uint8_t pod_group_type;
};
DF_HD void info_clear(InfoVal& v) {
  v.region_id = v.az_id = v.host_id = v.device_type = v.device_id = v.pod_node_id = v.pod_ns_id =
      v.pod_group_id = v.pod_id = v.pod_cluster_id = v.subnet_id = 0;
  v.l2_epc = 0;
  v.pod_group_type = 0;
}
DF_HD void info_load(const DevDict& d, uint32_t idx, InfoVal& v) {
  const DInfo& i = d.infos[idx];
  v.region_id = i.region_id;
  v.az_id = i.az_id;
  v.host_id = i.host_id;
  v.device_type = i.device_type;
  v.device_id = i.device_id;
  v.pod_node_id = i.pod_node_id;
  v.pod_ns_id = i.pod_ns_id;
  v.pod_group_id = i.pod_group_id;
  v.pod_id = i.pod_id;
  v.pod_cluster_id = i.pod_cluster_id;
  v.subnet_id = i.subnet_id;
  v.l2_epc = i.l2_epc;
  v.pod_group_type = i.pod_group_type;
}

// queryEpcIDBaseInfo :686-700
DF_HD bool query_epc_base(const DevDict& d, int32_t epc, uint32_t& region) {
  if (epc == EPC_UNKNOWN) {
    region = d.region_id;
    return true;
  }
  return t64_find(d.t64, d.t64_mask, TID_EPC_BASE, (uint64_t)(uint32_t)epc, region);
}

// QueryIPV4Infos :292-308 = exact (epc,ip) -> first containing CIDR of the epc
// (longest prefix first, :821-840) -> epc base region.
DF_KG_FN bool query_ip4(const DevDict& d, int32_t epc, uint32_t ip, InfoVal& out) {
  if (epc == EPC_FROM_INTERNET) return false;
  uint32_t idx;
  if (t64_find(d.t64, d.t64_mask, TID_IP4, (uint64_t)(int64_t)epc << 32 | (uint64_t)ip, idx)) {
    info_load(d, idx, out);
    return true;
  }
  uint32_t r;
  if (t64_find(d.t64, d.t64_mask, TID_CIDR4, (uint64_t)(uint32_t)epc, r)) {
    DRange rg = d.cidr4_ranges[r];
    for (uint32_t i = 0; i < rg.count; i++) {
      const DCidr4& c = d.cidr4[rg.start + i];
      if ((ip & c.mask) == c.net) {
        info_clear(out);
        out.subnet_id = c.subnet_id;
        out.region_id = c.region_id;
        out.az_id = c.az_id;
        return true;
      }
    }
  }
  uint32_t region;
  if (!query_epc_base(d, epc, region)) return false;
  info_clear(out);
  out.region_id = region;
  return true;
}

DF_HD bool ip6_is_v4mapped(const uint8_t* ip) {
  for (int i = 0; i < 10; i++)
    if (ip[i]) return false;
  return ip[10] == 0xff && ip[11] == 0xff;
}

// QueryIPV6Infos :310-326.  `ip` points at 16 bytes (zero padded when the wire
// carried fewer, like copy(key[4:], ipv6)); iplen is the wire length.
DF_KG_FN bool query_ip6(const DevDict& d, int32_t epc, const uint8_t* ip, uint32_t iplen, InfoVal& out) {
  if (epc == EPC_FROM_INTERNET) return false;
  uint32_t idx;
  if (t6_find(d.t6, d.t6_mask, T6_IP6, ip, epc, 0, 0, idx)) {
    info_load(d, idx, out);
    return true;
  }
  uint32_t r;
  // net.IPNet.Contains: only a 16-byte, non-v4-mapped address can match a v6 net
  if (iplen == 16 && !ip6_is_v4mapped(ip) &&
      t64_find(d.t64, d.t64_mask, TID_CIDR6, (uint64_t)(uint32_t)epc, r)) {
    DRange rg = d.cidr6_ranges[r];
    for (uint32_t i = 0; i < rg.count; i++) {
      const DCidr6& c = d.cidr6[rg.start + i];
      uint32_t full = c.plen >> 3, rem = c.plen & 7;
      bool m = true;
      for (uint32_t b = 0; b < full; b++) m &= ip[b] == c.net[b];
      if (m && rem) {
        uint8_t mk = (uint8_t)(0xff << (8 - rem));
        m = (ip[full] & mk) == (c.net[full] & mk);
      }
      if (m) {
        info_clear(out);
        out.subnet_id = c.subnet_id;
        out.region_id = c.region_id;
        out.az_id = c.az_id;
        return true;
      }
    }
  }
  uint32_t region;
  if (!query_epc_base(d, epc, region)) return false;
  info_clear(out);
  out.region_id = region;
  return true;
}

// queryMacInfo :613-623 (DupCount > 0 => nil)
DF_KG_FN bool query_mac(const DevDict& d, uint64_t l3_epc_mac, InfoVal& out) {
  uint32_t idx;
  if (!t64_find(d.t64, d.t64_mask, TID_MAC, l3_epc_mac, idx)) return false;
  if (d.infos[idx].dup) return false;
  info_load(d, idx, out);
  return true;
}
DF_KG_FN bool query_pod(const DevDict& d, uint32_t pod_id, InfoVal& out) {  // :376-382
  uint32_t idx;
  if (!t64_find(d.t64, d.t64_mask, TID_POD, pod_id, idx)) return false;
  info_load(d, idx, out);
  return true;
}

}  // namespace dfgpu
