/*
 * synth.cc — seeded synthetic workload generator: a controller
 * dictionary snapshot (trident.PlatformData / Groups / VtapIp / PodIp shaped, see
 * include/dfgpu.h) and agent frames carrying flow_log.proto / metric.proto
 * records, framed like agent/src/sender/uniform_sender.rs:135-196 writes them
 * ({u32 LE len, pb}* per frame; the 19 header bytes are consumed by the receiver
 * and surface as dfgpu_frame{vtap,org,team}).
 *
 * This is bench/test INPUT tooling, not part of the hot path.  The encoder
 * follows proto3 rules (zero scalars omitted, packed repeated scalars) like the
 * agent's prost encoder; `flags` adds wire-format edge cases for the parity tests.
 */
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <set>
#include <string>
#include <unordered_set>
#include <vector>

#include "../../include/dfgpu.h"

namespace {

struct Rng {
  uint64_t s;
  explicit Rng(uint64_t seed) : s(seed * 0x9E3779B97F4A7C15ull + 0xD1B54A32D192ED03ull) { next(); }
  uint64_t next() {  // splitmix64
    uint64_t z = (s += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
  }
  uint32_t below(uint32_t n) { return n ? (uint32_t)(((next() >> 32) * (uint64_t)n) >> 32) : 0; }
  double unit() { return (double)(next() >> 11) * (1.0 / 9007199254740992.0); }
  bool chance(double p) { return unit() < p; }
  // Zipf-like skew over [0,n): index = floor(n^u) - 1
  uint32_t skew(uint32_t n) {
    if (n <= 1) return 0;
    double v = std::exp(unit() * std::log((double)n));
    uint32_t i = (uint32_t)v;
    if (i >= 1) i -= 1;
    return i >= n ? n - 1 : i;
  }
  // log-uniform in [0, 2^bits)
  uint64_t logu(int bits) {
    int b = (int)below((uint32_t)bits + 1);
    if (b == 0) return 0;
    uint64_t hi = b >= 64 ? ~0ull : ((1ull << b) - 1);
    return next() & hi;
  }
};

// ---- protobuf writer ----------------------------------------------------------
struct PB {
  std::vector<uint8_t> b;
  void varint(uint64_t v) {
    while (v >= 0x80) {
      b.push_back((uint8_t)(v | 0x80));
      v >>= 7;
    }
    b.push_back((uint8_t)v);
  }
  void tag(uint32_t f, uint32_t wt) { varint((uint64_t)f << 3 | wt); }
  void u(uint32_t f, uint64_t v) {  // proto3: zero omitted
    if (v) {
      tag(f, 0);
      varint(v);
    }
  }
  void u_always(uint32_t f, uint64_t v) {
    tag(f, 0);
    varint(v);
  }
  void i32(uint32_t f, int32_t v) {  // int32: sign-extended to 64 bits on the wire
    if (v) {
      tag(f, 0);
      varint((uint64_t)(int64_t)v);
    }
  }
  void bytes(uint32_t f, const void* p, size_t n) {
    if (n) {
      tag(f, 2);
      varint(n);
      const uint8_t* q = (const uint8_t*)p;
      b.insert(b.end(), q, q + n);
    }
  }
  void str(uint32_t f, const std::string& s) { bytes(f, s.data(), s.size()); }
  void str_always(uint32_t f, const std::string& s) {
    tag(f, 2);
    varint(s.size());
    b.insert(b.end(), s.begin(), s.end());
  }
  void msg(uint32_t f, const PB& m) {  // Some(msg) is emitted even when empty
    tag(f, 2);
    varint(m.b.size());
    b.insert(b.end(), m.b.begin(), m.b.end());
  }
  void packed_u(uint32_t f, const std::vector<uint64_t>& v, bool packed) {
    if (v.empty()) return;
    if (packed) {
      PB t;
      for (auto x : v) t.varint(x);
      tag(f, 2);
      varint(t.b.size());
      b.insert(b.end(), t.b.begin(), t.b.end());
    } else {
      for (auto x : v) u_always(f, x);
    }
  }
  void packed_dbl(uint32_t f, const std::vector<double>& v, bool packed) {
    if (v.empty()) return;
    if (packed) {
      tag(f, 2);
      varint(v.size() * 8);
      for (auto d : v) {
        uint8_t t[8];
        memcpy(t, &d, 8);
        b.insert(b.end(), t, t + 8);
      }
    } else {
      for (auto d : v) {
        tag(f, 1);
        uint8_t t[8];
        memcpy(t, &d, 8);
        b.insert(b.end(), t, t + 8);
      }
    }
  }
};

enum : uint32_t {
  F_UNKNOWN_FIELDS = 1,   // sprinkle unknown fields (all wire types incl. groups)
  F_UNPACKED = 2,         // repeated scalars in the non-packed encoding
  F_DUP_FIELDS = 4,       // a scalar twice (last one wins) / split sub-messages (merge)
  F_OMIT_SUBMSG = 8,      // drop optional sub-messages entirely
  F_CORRUPT = 16,         // truncate / damage ~1 record per 50 frames
  F_WIDE_VARINT = 32,     // over-long varints, uint32 fields carrying >32-bit values
  F_HIGH_CARD = 64,       // L7: attribute values from a 2^17 space x 8 names (~1 M (name, value) pairs): the
                          // high-cardinality flow_tag workload of BASELINE config 5
};

struct Endpoint {
  int32_t epc;
  bool v6;
  uint32_t ip4;
  uint8_t ip6[16];
  uint64_t mac;
  uint32_t pod_id, pod_node_id, pod_cluster_id, pod_group_id, pod_ns_id, gpid, pid, vtap;
};

}  // namespace

struct synth_ctx {
  uint64_t seed;
  uint32_t P;
  std::vector<dfgpu_interface> interfaces;
  std::vector<dfgpu_ip_resource> ip_resources;
  std::vector<dfgpu_cidr> cidrs;
  std::vector<dfgpu_gprocess> gprocesses;
  std::vector<dfgpu_vtap_ip> vtap_ips;
  std::vector<dfgpu_pod_ip> pod_ips;
  std::vector<dfgpu_service> services;
  std::vector<uint32_t> u32_pool;
  std::vector<dfgpu_ip> ip_pool;
  dfgpu_dict_snapshot snap;
  std::vector<Endpoint> endpoints;
  std::vector<uint32_t> svc_ports;
  uint32_t n_vtap, n_epc;
  std::vector<dfgpu_geo_entry> geo;
  std::vector<std::string> geo_names;
  std::vector<const char*> geo_name_ptrs;
};

struct synth_batch {
  std::vector<uint8_t> data;
  std::vector<dfgpu_frame> frames;
  uint64_t nrecords = 0;
};

struct synth_frames {
  const uint8_t* data;
  uint64_t nbytes;
  const dfgpu_frame* frames;
  uint32_t nframes;
  uint32_t pad;
  uint64_t nrecords;
};

static dfgpu_ip mk_ip4(uint32_t ip) {
  dfgpu_ip r;
  memset(&r, 0, sizeof r);
  r.valid = 1;
  r.ip4 = ip;
  return r;
}
static dfgpu_ip mk_ip6(const uint8_t* ip) {
  dfgpu_ip r;
  memset(&r, 0, sizeof r);
  r.valid = 1;
  r.is_v6 = 1;
  memcpy(r.ip6, ip, 16);
  return r;
}

static void build_dict(synth_ctx* c) {
  Rng r(c->seed ^ 0x5eed0001);
  uint32_t P = c->P;
  c->n_epc = std::max<uint32_t>(4, std::min<uint32_t>(2000, P / 50));
  c->n_vtap = std::max<uint32_t>(8, std::min<uint32_t>(2000, P / 50));
  uint32_t n_node = std::max<uint32_t>(2, P / 30), n_group = std::max<uint32_t>(2, P / 5);
  std::set<std::pair<int32_t, uint32_t>> used_ip;
  c->interfaces.reserve(P);
  c->ip_resources.reserve(P + P / 20);
  uint64_t prev_mac = 0;
  int32_t prev_epc = 0;
  for (uint32_t i = 0; i < P; i++) {
    dfgpu_interface f;
    memset(&f, 0, sizeof f);
    uint32_t kind = r.below(100);
    f.device_type = kind < 70 ? 10 : kind < 85 ? 1 : kind < 95 ? 14 : kind < 98 ? 11 : 6;
    f.device_id = i + 1;
    f.epc_id = 1 + r.below(c->n_epc);
    if (r.chance(0.01)) f.epc_id = 0;  // 0 -> EPC_FROM_INTERNET (-2)
    f.if_type = r.chance(0.03) ? 3 : 4;  // 3 = WAN
    f.launch_server_id = 1 + r.below(std::max<uint32_t>(2, P / 40));
    f.region_id = 1 + r.below(3);
    f.az_id = 1 + r.below(6);
    f.vtap_id = 1 + r.below(c->n_vtap);
    f.pod_cluster_id = 1 + r.below(20);
    f.pod_node_id = 1 + r.below(n_node);
    if (f.device_type == 10) {
      f.pod_id = i + 1;
      f.pod_ns_id = 1 + r.below(500);
      f.pod_group_id = 1 + r.below(n_group);
      f.pod_group_type = (uint32_t[]){130, 131, 132, 133}[r.below(4)];
    } else if (f.device_type != 14) {
      f.pod_node_id = 0;
      f.pod_cluster_id = r.chance(0.3) ? f.pod_cluster_id : 0;
    }
    f.mac = 0x020000000000ull | ((uint64_t)(i + 1) * 2654435761ull & 0xffffffffffull);
    if (r.chance(0.01) && prev_mac) {  // duplicate (mac, epc) => DupCount path
      f.mac = prev_mac;
      f.epc_id = prev_epc == -2 ? 0 : (uint32_t)prev_epc;
    }
    if (r.chance(0.02)) f.mac = 0;
    prev_mac = f.mac;
    prev_epc = f.epc_id == 0 ? -2 : (int32_t)f.epc_id;
    f.ip_res_off = (uint32_t)c->ip_resources.size();
    uint32_t nip = r.chance(0.05) ? 2 : 1;
    int32_t epc = f.epc_id == 0 ? -2 : (int32_t)f.epc_id;
    for (uint32_t k = 0; k < nip; k++) {
      dfgpu_ip_resource ir;
      memset(&ir, 0, sizeof ir);
      ir.subnet_id = 1 + r.below(4000);
      Endpoint e;
      memset(&e, 0, sizeof e);
      e.epc = epc;
      if (r.chance(0.10)) {
        uint8_t ip6[16] = {0xfd, 0x00};
        uint64_t a = r.next(), b2 = r.next();
        memcpy(ip6 + 2, &a, 6);
        memcpy(ip6 + 8, &b2, 8);
        ir.ip = mk_ip6(ip6);
        e.v6 = true;
        memcpy(e.ip6, ip6, 16);
      } else {
        uint32_t ip;
        do {
          ip = 0x0A000000u | (r.below(1 << 24));
          if ((ip & 0xff) == 0) ip |= 1;
        } while (!used_ip.insert({epc, ip}).second);
        ir.ip = mk_ip4(ip);
        e.ip4 = ip;
      }
      c->ip_resources.push_back(ir);
      e.mac = f.mac;
      e.pod_id = f.pod_id;
      e.pod_node_id = f.pod_node_id;
      e.pod_cluster_id = f.pod_cluster_id;
      e.pod_group_id = f.pod_group_id;
      e.pod_ns_id = f.pod_ns_id;
      e.vtap = f.vtap_id;
      c->endpoints.push_back(e);
    }
    f.ip_res_cnt = nip;
    c->interfaces.push_back(f);
  }
  // gprocesses: attached to pod endpoints
  uint32_t n_gp = std::max<uint32_t>(16, P / 2);
  for (uint32_t i = 0; i < n_gp; i++) {
    Endpoint& e = c->endpoints[r.below((uint32_t)c->endpoints.size())];
    dfgpu_gprocess g;
    g.gprocess_id = i + 1;
    g.vtap_id = e.vtap;
    g.pod_id = r.chance(0.85) ? e.pod_id : 0;
    g.pid = r.chance(0.9) ? 1000 + r.below(60000) : 0;
    if (r.chance(0.01)) g.vtap_id = 0;
    c->gprocesses.push_back(g);
    if (e.gpid == 0) {
      e.gpid = g.gprocess_id;
      e.pid = g.pid;
    }
  }
  // cidrs
  uint32_t n_cidr = std::max<uint32_t>(8, std::min<uint32_t>(2000, P / 50));
  std::set<std::tuple<int32_t, uint32_t, uint32_t>> used_cidr;
  std::set<std::pair<uint32_t, uint32_t>> used_wan;
  for (uint32_t i = 0; i < n_cidr; i++) {
    dfgpu_cidr ci;
    memset(&ci, 0, sizeof ci);
    ci.epc_id = (int32_t)(1 + r.below(c->n_epc));
    if (r.chance(0.02)) ci.epc_id = 0;
    ci.type = r.chance(0.05) ? 1 : 2;
    ci.subnet_id = 5000 + i;
    ci.region_id = 1 + r.below(3);
    ci.az_id = 1 + r.below(6);
    if (r.chance(0.08)) {
      uint8_t ip6[16] = {0xfd, 0x00};
      uint64_t a = r.next();
      int plen = 32 + 8 * (int)r.below(5);  // /32../64
      memcpy(ip6 + 2, &a, 6);
      for (int b = plen / 8; b < 16; b++) ip6[b] = 0;
      ci.prefix = mk_ip6(ip6);
      ci.prefix_len = (uint32_t)plen;
      uint32_t h = (uint32_t)(a ^ (a >> 32));
      if (!used_cidr.insert({ci.epc_id, h, (uint32_t)plen + 1000}).second) continue;
    } else {
      uint32_t plen = 16 + r.below(13);
      uint32_t mask = 0xffffffffu << (32 - plen);
      uint32_t net = (0x0A000000u | r.below(1 << 24)) & mask;
      if (r.chance(0.1)) net = (0xC0A80000u | r.below(1 << 16)) & mask;
      ci.prefix = mk_ip4(net);
      ci.prefix_len = plen;
      int32_t e2 = ci.epc_id == 0 ? -2 : ci.epc_id;
      if (!used_cidr.insert({e2, net, plen}).second) continue;
      if (ci.type == 1 && !used_wan.insert({net, plen}).second) continue;
    }
    c->cidrs.push_back(ci);
  }
  // vtaps: ip of some endpoint (so the agent lookup can hit)
  for (uint32_t v = 1; v <= c->n_vtap; v++) {
    dfgpu_vtap_ip vi;
    memset(&vi, 0, sizeof vi);
    vi.vtap_id = v;
    const Endpoint& e = c->endpoints[r.below((uint32_t)c->endpoints.size())];
    vi.epc_id = e.epc == -2 ? 0 : (uint32_t)e.epc;
    vi.ip = e.v6 ? mk_ip6(e.ip6) : mk_ip4(e.ip4);
    if (r.chance(0.05)) vi.ip = mk_ip4(0x0A000000u | r.below(1 << 24));  // not a known endpoint
    vi.pod_cluster_id = e.pod_cluster_id;
    vi.team_id = 1 + r.below(3);
    if (!r.chance(0.02)) c->vtap_ips.push_back(vi);  // a few vtaps unknown to the controller
  }
  // pod_ips: every 10th interface pod repeated (ignored: already known) + extra pods
  uint32_t n_extra = std::max<uint32_t>(4, P / 50);
  for (uint32_t i = 0; i < n_extra; i++) {
    dfgpu_pod_ip p;
    memset(&p, 0, sizeof p);
    p.pod_id = P + 1 + i;
    const Endpoint& e = c->endpoints[r.below((uint32_t)c->endpoints.size())];
    p.epc_id = e.epc == -2 ? 0 : (uint32_t)e.epc;
    if (r.chance(0.7)) p.ip = e.v6 ? mk_ip6(e.ip6) : mk_ip4(e.ip4);  // host-network pod: node's ip
    else if (r.chance(0.5)) p.pod_node_ip = mk_ip4(0x0A000000u | r.below(1 << 24));
    p.pod_cluster_id = 1 + r.below(20);
    p.pod_ns_id = 1 + r.below(500);
    p.pod_group_id = 1 + r.below(n_group);
    p.pod_group_type = 130;
    c->pod_ips.push_back(p);
    Endpoint x = e;
    x.pod_id = p.pod_id;
    x.mac = 0;
    x.gpid = 0;
    c->endpoints.push_back(x);
  }
  for (uint32_t i = 0; i < P; i += 10) {
    const dfgpu_interface& f = c->interfaces[i];
    if (!f.pod_id) continue;
    dfgpu_pod_ip p;
    memset(&p, 0, sizeof p);
    p.pod_id = f.pod_id;
    p.epc_id = f.epc_id;
    p.ip = c->ip_resources[f.ip_res_off].ip;
    p.pod_cluster_id = f.pod_cluster_id;
    c->pod_ips.push_back(p);
  }
  // services
  uint32_t n_svc = std::max<uint32_t>(16, P / 20);
  for (uint32_t i = 0; i < 64; i++) c->svc_ports.push_back(i < 8 ? (uint32_t[]){80, 443, 8080, 3306, 6379, 53, 9092, 8443}[i] : 1024 + r.below(30000));
  auto add_ports = [&](dfgpu_service& s, uint32_t n) {
    s.server_ports_off = (uint32_t)c->u32_pool.size();
    for (uint32_t k = 0; k < n; k++) c->u32_pool.push_back(c->svc_ports[r.below((uint32_t)c->svc_ports.size())]);
    s.server_ports_cnt = n;
  };
  for (uint32_t i = 0; i < n_svc; i++) {
    dfgpu_service s;
    memset(&s, 0, sizeof s);
    s.id = i + 1;
    s.protocol = r.below(3);
    uint32_t kind = r.below(100);
    const Endpoint& e = c->endpoints[r.below((uint32_t)c->endpoints.size())];
    if (kind < 40) {
      s.type = 2;  // POD_SERVICE_POD_GROUP
      s.pod_group_ids_off = (uint32_t)c->u32_pool.size();
      c->u32_pool.push_back(e.pod_group_id ? e.pod_group_id : 1 + r.below(n_group));
      s.pod_group_ids_cnt = 1;
      add_ports(s, 1 + r.below(3));
    } else if (kind < 55) {
      s.type = 1;  // POD_SERVICE_NODE
      s.pod_cluster_id = 1 + r.below(20);
      add_ports(s, 1 + r.below(2));
    } else if (kind < 80) {
      s.type = 3;  // POD_SERVICE_IP
      s.epc_id = e.epc > 0 ? (uint32_t)e.epc : 1;
      s.ips_off = (uint32_t)c->ip_pool.size();
      c->ip_pool.push_back(e.v6 ? mk_ip6(e.ip6) : mk_ip4(e.ip4));
      s.ips_cnt = 1;
      add_ports(s, 1 + r.below(2));
    } else {
      s.type = 7;  // CUSTOM_SERVICE
      uint32_t sub = r.below(8);
      s.epc_id = e.epc > 0 ? (uint32_t)e.epc : 0;
      auto one = [&](uint32_t& off, uint32_t& cnt, uint32_t v) {
        off = (uint32_t)c->u32_pool.size();
        c->u32_pool.push_back(v);
        cnt = 1;
      };
      switch (sub) {
        case 0: one(s.pod_service_ids_off, s.pod_service_ids_cnt, 1 + r.below(n_svc)); break;
        case 1: one(s.pod_group_ids_off, s.pod_group_ids_cnt, 1 + r.below(n_group)); break;
        case 2: one(s.pod_ids_off, s.pod_ids_cnt, e.pod_id ? e.pod_id : 1); break;
        case 3: one(s.chost_ids_off, s.chost_ids_cnt, 1 + r.below(P)); break;
        case 4: one(s.host_ids_off, s.host_ids_cnt, 1 + r.below(P)); break;
        case 5:
          s.pod_cluster_id = 1 + r.below(20);
          if (r.chance(0.5)) add_ports(s, 1 + r.below(2));
          break;
        default:
          s.ips_off = (uint32_t)c->ip_pool.size();
          c->ip_pool.push_back(e.v6 ? mk_ip6(e.ip6) : mk_ip4(e.ip4));
          s.ips_cnt = 1;
          if (r.chance(0.3)) s.epc_id = 0;
          if (r.chance(0.6)) add_ports(s, 1 + r.below(2));
          break;
      }
    }
    c->services.push_back(s);
  }
  memset(&c->snap, 0, sizeof c->snap);
  c->snap.version = c->seed;
  c->snap.region_id = 1;
  c->snap.analyzer_id = 7;
  c->snap.interfaces = c->interfaces.data();
  c->snap.n_interfaces = c->interfaces.size();
  c->snap.ip_resources = c->ip_resources.data();
  c->snap.n_ip_resources = c->ip_resources.size();
  c->snap.cidrs = c->cidrs.data();
  c->snap.n_cidrs = c->cidrs.size();
  c->snap.gprocesses = c->gprocesses.data();
  c->snap.n_gprocesses = c->gprocesses.size();
  c->snap.vtap_ips = c->vtap_ips.data();
  c->snap.n_vtap_ips = c->vtap_ips.size();
  c->snap.pod_ips = c->pod_ips.data();
  c->snap.n_pod_ips = c->pod_ips.size();
  c->snap.services = c->services.data();
  c->snap.n_services = c->services.size();
  c->snap.u32_pool = c->u32_pool.data();
  c->snap.n_u32_pool = c->u32_pool.size();
  c->snap.ip_pool = c->ip_pool.data();
  c->snap.n_ip_pool = c->ip_pool.size();

  // geo: disjoint public ranges, 34 regions (names[0] = unknown, ip_info.py:41)
  Rng g(c->seed ^ 0x9e0);
  c->geo_names.push_back("\xe6\x9c\xaa\xe7\x9f\xa5");
  for (int i = 1; i <= 34; i++) {
    char t[32];
    snprintf(t, sizeof t, "province-%02d", i);
    c->geo_names.push_back(t);
  }
  // /24-aligned blocks (like allocation data), plus a few unaligned odd-sized ranges
  uint32_t cur = 0x01000000u;
  uint32_t n_odd = 0;
  while (cur < 0xDF000000u && c->geo.size() < 6000) {
    cur += (g.below(1 << 10)) << 8;
    uint32_t len = 256u << g.below(11);
    if ((cur >> 24) == 10 || (cur >> 24) == 127) {
      cur = ((cur >> 24) + 1) << 24;
      continue;
    }
    dfgpu_geo_entry e;
    memset(&e, 0, sizeof e);
    e.ip_start = cur;
    e.ip_end = cur + len - 1;
    if (n_odd < 40 && g.chance(0.01)) {  // unaligned start/end inside the block
      e.ip_start += 1 + g.below(100);
      e.ip_end -= 1 + g.below(100);
      n_odd++;
    }
    e.region = (uint8_t)(g.chance(0.03) ? 0 : 1 + g.below(34));
    c->geo.push_back(e);
    cur += len;
  }
  for (auto& s : c->geo_names) c->geo_name_ptrs.push_back(s.c_str());
}

// ---- record generators ------------------------------------------------------------
namespace {

struct Side {
  int32_t epc;
  bool v6;
  uint32_t ip4;
  uint8_t ip6[16];
  uint64_t mac;
  uint32_t gpid, pod_id, pid;
};

void pick_side(synth_ctx* c, Rng& r, Side& s, bool want_v6, int mode /*0 known,1 internet,2 unknown*/,
               uint32_t hot_endpoints = 0 /*>0: known endpoints come from the first N only*/) {
  memset(&s, 0, sizeof s);
  s.v6 = want_v6;
  if (mode == 0) {
    for (int tries = 0; tries < 8; tries++) {
      uint32_t pool = hot_endpoints ? std::min<uint32_t>(hot_endpoints, (uint32_t)c->endpoints.size())
                                       : (uint32_t)c->endpoints.size();
      const Endpoint& e = c->endpoints[r.skew(pool)];
      if (e.v6 != want_v6) continue;
      s.epc = e.epc;
      s.ip4 = e.ip4;
      memcpy(s.ip6, e.ip6, 16);
      s.mac = e.mac;
      s.gpid = e.gpid;
      s.pod_id = e.pod_id;
      s.pid = e.pid;
      if (r.chance(0.05)) s.ip4 ^= 1 + r.below(255);  // same subnet, unknown host -> CIDR path
      if (r.chance(0.05)) s.epc = (int32_t)(1 + r.below(hot_endpoints ? 10 : c->n_epc));  // wrong epc
      return;
    }
    mode = 2;
  }
  if (mode == 1) {
    s.epc = -2;
    if (want_v6) {
      uint64_t a = r.next(), b = r.next();
      s.ip6[0] = 0x20;
      s.ip6[1] = 0x01;
      memcpy(s.ip6 + 2, &a, 6);
      memcpy(s.ip6 + 8, &b, 8);
    } else {
      s.ip4 = c->geo.empty() ? 0x08080808u : c->geo[r.below((uint32_t)c->geo.size())].ip_start + r.below(200);
      if (r.chance(0.3)) s.ip4 = 0x01000000u + (uint32_t)(r.next() % 0xDE000000u);
    }
    s.mac = r.chance(0.5) ? 0x0200000000ffull + r.below(1000) : 0;
    return;
  }
  // unknown: epc 0 / -1 / random, special addresses
  uint32_t k = r.below(6);
  s.epc = k == 0 ? 0 : k == 1 ? -1 : (int32_t)(1 + r.below(hot_endpoints ? 10 : c->n_epc));
  if (want_v6) {
    uint32_t sp = r.below(6);
    if (sp == 0) s.ip6[15] = 1;                                     // ::1
    else if (sp == 1) { s.ip6[0] = 0xfe; s.ip6[1] = 0x80; s.ip6[15] = 7; }  // link local
    else if (sp == 2) { s.ip6[0] = 0xff; s.ip6[1] = 0x02; s.ip6[15] = 1; }  // multicast
    else { s.ip6[0] = 0xfd; uint64_t a = r.next(); memcpy(s.ip6 + 8, &a, 8); }
  } else {
    uint32_t sp = r.below(8);
    if (sp == 0) s.ip4 = 0x7F000001u;
    else if (sp == 1) s.ip4 = 0xA9FE0000u | r.below(65536);
    else if (sp == 2) s.ip4 = 0xE0000000u | r.below(1 << 20);
    else if (sp == 3) s.ip4 = 0xFFFFFFFFu;
    else if (sp == 4) s.ip4 = 0;
    else s.ip4 = 0x0A000000u | r.below(1 << 24);
  }
  s.mac = r.chance(0.5) ? 0x020000abc000ull + r.below(4096) : 0;
  s.gpid = r.chance(0.2) ? 1 + r.below(100000) : 0;
}

const uint32_t kTapSides[] = {0, 1, 2, 4, 9, 10, 17, 18, 25, 26, 33, 34, 41, 42, 49, 50, 48};
const uint32_t kCloseTypes[] = {0, 1, 2, 3, 5, 7, 8, 9, 10, 11, 13, 15, 17, 18, 19, 20};

void unknown_fields(PB& pb, Rng& r) {
  switch (r.below(5)) {
    case 0: pb.u_always(200 + r.below(50), r.next()); break;
    case 1: pb.tag(300, 1); for (int i = 0; i < 8; i++) pb.b.push_back((uint8_t)r.next()); break;
    case 2: pb.str_always(301, "ignored-bytes"); break;
    case 3: pb.tag(302, 5); for (int i = 0; i < 4; i++) pb.b.push_back((uint8_t)r.next()); break;
    case 4: pb.tag(303, 3); pb.u_always(1, 5); pb.tag(303, 4); break;  // group
  }
}

void gen_peer(PB& pb, Rng& r, const Side& s, uint32_t flags, bool tcp) {
  pb.u(1, r.logu(40));
  pb.u(2, r.logu(40));
  pb.u(3, r.logu(40));
  pb.u(4, r.logu(32));
  pb.u(5, r.logu(44));
  pb.u(6, r.logu(36));
  pb.u(7, r.next() >> 4);
  pb.u(8, r.next() >> 4);
  if (tcp) pb.u(9, r.below(256) | (r.chance(0.05) ? 0x10000 : 0));
  int32_t epc = s.epc;
  if (r.chance(0.01)) epc = -(int32_t)(3 + r.below(40000));  // < -2: MarshalInt32WithSpecialID path
  pb.i32(10, epc);
  pb.u(11, r.below(2));
  pb.u(12, r.below(2));
  pb.u(13, r.below(2));
  pb.u(14, r.below(2));
  pb.u(15, r.chance(0.03) ? 1 : 0);
  pb.u(16, r.chance(0.03) ? 1 : 0);
  if (r.chance(0.05)) {
    pb.u(20, 0x0A000000u | r.below(1 << 24));
    pb.u(21, 1024 + r.below(60000));
  }
  pb.u(22, s.gpid);
  if (flags & F_UNKNOWN_FIELDS)
    if (r.chance(0.1)) unknown_fields(pb, r);
}

void gen_l4(synth_ctx* c, Rng& r, uint32_t vtap, uint64_t base_ns, uint32_t flags, PB& out) {
  bool v6 = r.chance(0.10);
  Side a, b;
  uint32_t m = r.below(100);
  pick_side(c, r, a, v6, m < 80 ? 0 : m < 88 ? 1 : m < 95 ? 0 : 2);
  pick_side(c, r, b, v6, m < 80 ? 0 : m < 88 ? 0 : m < 95 ? 1 : 2);
  bool tcp = r.chance(0.70);
  uint32_t proto = tcp ? 6 : (r.chance(0.8) ? 17 : 1);
  uint32_t tap_side = kTapSides[r.skew(17)];
  if (r.chance(0.25)) tap_side = r.chance(0.5) ? 41 : 42;
  if (r.chance(0.08)) tap_side = 4;

  PB fk;
  fk.u(1, r.chance(0.97) ? vtap : 1 + r.below(c->n_vtap));
  fk.u(2, r.chance(0.9) ? 3 : r.below(256));
  uint64_t tap_port = (uint64_t)r.below(1u << 31) | (uint64_t)r.below(16) << 32 |
                      (uint64_t)(uint32_t[]){0, 2, 4, 6}[r.below(4)] << 36 | (uint64_t)r.below(9) << 60;
  fk.u(3, tap_port);
  fk.u(4, a.mac);
  fk.u(5, b.mac);
  if (!v6 || r.chance(0.2)) {
    fk.u(6, a.ip4);
    fk.u(7, b.ip4);
  }
  if (v6) {
    fk.bytes(8, a.ip6, 16);
    fk.bytes(9, b.ip6, 16);
  }
  uint32_t sport = 1024 + r.below(60000);
  uint32_t dport = r.chance(0.7) ? c->svc_ports[r.below((uint32_t)c->svc_ports.size())] : r.below(65536);
  fk.u(10, sport | ((flags & F_WIDE_VARINT) && r.chance(0.05) ? 0x10000u : 0));
  fk.u(11, dport);
  fk.u(12, proto);
  if (flags & F_UNKNOWN_FIELDS)
    if (r.chance(0.1)) unknown_fields(fk, r);

  PB ps, pd, tun;
  gen_peer(ps, r, a, flags, tcp);
  gen_peer(pd, r, b, flags, tcp);
  bool tunnel = r.chance(0.03);
  if (tunnel) {
    tun.u(1, r.next() & 0xffffffffu);
    tun.u(2, r.next() & 0xffffffffu);
    tun.u(3, r.next() & 0xffffffffu);
    tun.u(4, r.next() & 0xffffffffu);
    tun.u(5, r.next() & 0xffffffffu);
    tun.u(6, r.next() & 0xffffffffu);
    tun.u(7, r.next() & 0xffffffffu);
    tun.u(8, r.next() & 0xffffffffu);
    tun.u(9, r.below(1 << 24));
    tun.u(10, r.below(1 << 24));
    tun.u(11, 1 + r.below(5));
    tun.u(12, 1 + r.below(2));
  } else if (r.chance(0.02)) {
    tun.u(1, 12345);  // tunnel fields present but type NONE: must be ignored
    tun.u(9, 77);
  }

  PB f;
  bool dup = (flags & F_DUP_FIELDS) && r.chance(0.2);
  if (dup) {  // split flow_key in two chunks: gogo merges
    PB k1, k2;
    size_t half = fk.b.size();  // cannot split mid-field; emit an extra partial message first
    (void)half;
    k1.u(1, 4242);  // overwritten by the complete key that follows
    k1.u(12, 99);
    f.msg(1, k1);
  }
  f.msg(1, fk);
  f.msg(2, ps);
  f.msg(3, pd);
  if (!((flags & F_OMIT_SUBMSG) && !tunnel && r.chance(0.5))) f.msg(4, tun);
  f.u(5, r.next() >> 1);
  uint64_t end_ns = base_ns + (uint64_t)r.below(120) * 1000000000ull + r.below(1000000000u);
  uint64_t dur = r.logu(36);
  f.u(6, end_ns - std::min(dur, end_ns));
  f.u(7, end_ns);
  f.u(8, dur);
  if (r.chance(0.1)) f.u(10, r.below(4096) | ((flags & F_WIDE_VARINT) && r.chance(0.3) ? 0x10000u : 0));
  f.u(11, v6 ? 0x86DD : 0x0800);
  bool perf = tcp ? r.chance(0.95) : r.chance(0.3);
  f.u(12, perf ? 1 : (r.chance(0.02) ? 2 : 0));
  if (perf || r.chance(0.05)) {
    PB t, l7, p;
    t.u(1, r.logu(24)); t.u(2, r.logu(24)); t.u(3, r.logu(24)); t.u(4, r.logu(24)); t.u(5, r.logu(24));
    t.u(8, r.logu(30)); t.u(9, r.logu(30)); t.u(12, r.logu(16)); t.u(13, r.logu(16));
    if (!((flags & F_OMIT_SUBMSG) && r.chance(0.3))) {
      PB c1, c2;
      c1.u(1, r.logu(12)); c1.u(2, r.logu(10)); c1.u(3, r.logu(10));
      c2.u(1, r.logu(12)); c2.u(2, r.logu(10)); c2.u(3, r.logu(10));
      t.msg(14, c1);
      t.msg(15, c2);
    }
    t.u(16, r.logu(12)); t.u(17, r.below(5)); t.u(18, r.below(5)); t.u(19, r.logu(24));
    t.u(20, r.logu(30)); t.u(21, r.logu(16)); t.u(22, r.below(4));
    l7.u(1, r.logu(20)); l7.u(2, r.logu(20)); l7.u(3, r.logu(10)); l7.u(4, r.logu(10)); l7.u(5, r.logu(8));
    l7.u(6, r.logu(20));
    l7.u(7, r.chance(0.05) ? (r.next() | (1ull << 60)) : r.logu(44));  // > 2^53 exercises the f64 cast
    l7.u(8, r.logu(24)); l7.u(9, r.logu(20));
    if (tcp) p.msg(1, t);
    p.msg(2, l7);
    p.u(3, tcp ? 1 : 2);
    p.u(4, (uint32_t[]){0, 20, 21, 40, 60, 80, 100, 120}[r.below(8)]);
    p.u(5, r.logu(8));
    f.msg(13, p);
  }
  f.u(14, kCloseTypes[r.below(16)] | ((flags & F_WIDE_VARINT) && r.chance(0.02) ? 0x100u : 0));
  f.u(15, (uint32_t[]){0, 0, 0, 1, 3, 4}[r.below(6)]);
  f.u(16, r.below(2));
  f.u(17, r.below(16));
  f.u(18, r.below(2));
  f.u(19, tap_side);
  if (dup) f.u_always(19, tap_side);
  if (tcp) {
    f.u(20, (uint32_t)r.next()); f.u(21, (uint32_t)r.next());
    if (r.chance(0.2)) { f.u(22, (uint32_t)r.next()); f.u(23, (uint32_t)r.next()); }
  }
  if (r.chance(0.02)) {
    std::vector<uint64_t> g;
    uint32_t n = 1 + r.below(4);
    for (uint32_t i = 0; i < n; i++) g.push_back(1 + r.below(70000));
    f.packed_u(24, g, !(flags & F_UNPACKED) || r.chance(0.5));
  }
  f.u(25, r.below(256));
  if (r.chance(0.15)) {
    char t[64];
    snprintf(t, sizeof t, "svc-%u.ns-%u.example.com", r.skew(5000), r.below(50));
    f.str(26, t);
  }
  if (r.chance(0.05)) {
    std::vector<uint64_t> g;
    uint32_t n = 1 + r.below(8);
    for (uint32_t i = 0; i < n; i++) g.push_back(r.chance(0.1) ? r.next() : r.logu(40));
    f.packed_u(27, g, !(flags & F_UNPACKED) || r.chance(0.5));
  }
  f.u(28, r.below(65536));
  if (flags & F_UNKNOWN_FIELDS)
    if (r.chance(0.15)) unknown_fields(f, r);
  out.msg(1, f);
  if ((flags & F_UNKNOWN_FIELDS) && r.chance(0.05)) unknown_fields(out, r);
}

std::string hexid(Rng& r, int n) {
  static const char* h = "0123456789abcdef";
  std::string s;
  for (int i = 0; i < n; i++) s.push_back(h[r.below(16)]);
  return s;
}

void gen_l7(synth_ctx* c, Rng& r, uint32_t vtap, uint64_t base_ns, uint32_t flags, PB& out) {
  bool v6 = r.chance(0.10);
  Side a, b;
  uint32_t m = r.below(100);
  // application traffic concentrates on a bounded set of service pods: keeps the
  // (vpc, pod_ns, name, value) flow_tag key space below the reference's 2^18 LRU
  // capacity, as a saturated production tag set would be
  pick_side(c, r, a, v6, m < 80 ? 0 : m < 88 ? 1 : m < 95 ? 0 : 2, 120);
  pick_side(c, r, b, v6, m < 80 ? 0 : m < 88 ? 0 : m < 95 ? 1 : 2, 120);
  uint32_t l7proto = r.chance(0.6) ? 20 : 21;
  if (flags & F_DUP_FIELDS)  // protocol variety only in the parity-test mode
    l7proto = (uint32_t[]){20, 21, 40, 43, 60, 61, 80, 100, 101, 120, 121, 127, 0, 99}[r.below(14)];
  uint32_t msg_type = (uint32_t[]){0, 1, 2, 2, 2}[r.below(5)];
  uint64_t end_ns = base_ns + (uint64_t)r.below(120) * 1000000000ull + r.below(1000000000u);
  uint64_t rrt = r.logu(32);
  bool ebpf = r.chance(0.6);
  uint32_t tap_side = ebpf ? (r.chance(0.5) ? 41 : 42) : kTapSides[r.skew(13)];

  PB head;
  head.u(1, l7proto);
  head.u(2, msg_type);
  head.u(5, rrt);
  PB base;
  base.u(1, end_ns - std::min<uint64_t>(rrt, end_ns));
  base.u(2, end_ns);
  base.u(3, r.next() >> 1);
  uint64_t tap_port = (uint64_t)r.below(1u << 31) | (uint64_t)r.below(16) << 32 |
                      (uint64_t)(uint32_t[]){0, 2, 4, 6}[r.below(4)] << 36 |
                      (uint64_t)(ebpf ? 7 : r.below(9)) << 60;
  base.u(4, tap_port);
  base.u(5, r.chance(0.97) ? vtap : 1 + r.below(c->n_vtap));
  base.u(6, 3);
  base.u(7, v6 ? 1 : 0);
  base.u(8, tap_side);
  base.msg(9, head);
  base.u(10, a.mac);
  base.u(11, b.mac);
  if (!v6 || r.chance(0.2)) {
    base.u(12, a.ip4);
    base.u(13, b.ip4);
  }
  if (v6) {
    base.bytes(14, a.ip6, 16);
    base.bytes(15, b.ip6, 16);
  }
  base.i32(16, a.epc);
  base.i32(17, b.epc);
  base.u(18, 1024 + r.below(60000));
  base.u(19, r.chance(0.7) ? c->svc_ports[r.below((uint32_t)c->svc_ports.size())] : r.below(65536));
  base.u(20, r.chance(0.9) ? 6 : 17);
  base.u(21, r.chance(0.02) ? 1 : 0);
  base.u(22, r.chance(0.02) ? 1 : 0);
  base.u(23, (uint32_t)r.next());
  base.u(24, (uint32_t)r.next());
  bool with_proc = ebpf && r.chance(0.8);
  if (with_proc) {
    base.u(25, a.pid ? a.pid : (r.chance(0.5) ? 1000 + r.below(60000) : 0));
    base.u(26, b.pid ? b.pid : (r.chance(0.5) ? 1000 + r.below(60000) : 0));
    base.str(27, (const char*[]){"nginx", "java", "python3", "envoy", "node", ""}[r.below(6)]);
    base.str(28, (const char*[]){"nginx", "java", "mysqld", "redis-server", "envoy", ""}[r.below(6)]);
    base.u(29, r.next() >> 8);
    base.u(30, r.next() >> 8);
    base.u(31, r.below(1 << 22));
    base.u(32, r.below(1 << 22));
    base.u(33, r.below(1 << 20));
    base.u(34, r.below(1 << 20));
    if (r.chance(0.2)) { base.u(39, r.next() >> 16); base.u(40, r.next() >> 16); }
  }
  bool tags = r.chance(0.6);
  base.u(35, tags && r.chance(0.8) ? a.gpid : 0);
  base.u(36, tags && r.chance(0.8) ? b.gpid : 0);
  base.u(41, tags && r.chance(0.7) ? a.pod_id : 0);
  base.u(42, tags && r.chance(0.7) ? b.pod_id : 0);
  if (r.chance(0.05)) base.u(43, 1 + r.below(5));
  if ((flags & F_UNKNOWN_FIELDS) && r.chance(0.1)) unknown_fields(base, r);

  PB req, resp, trace, ext;
  char t[160];
  static const char* methods[] = {"GET", "POST", "PUT", "DELETE", "HEAD", "PATCH"};
  if (msg_type != 1 || r.chance(0.2)) {
    req.str(1, l7proto == 100 ? (const char*[]){"Fetch", "Produce", "Metadata", "Heartbeat"}[r.below(4)] : methods[r.skew(6)]);
    snprintf(t, sizeof t, "svc-%u.ns-%u.example.com", r.skew(5000), r.below(50));
    req.str(2, t);
    snprintf(t, sizeof t, "/api/v%u/resource/%u/items/%u", 1 + r.below(3), r.skew(200000), r.below(1000));
    req.str(3, t);
    snprintf(t, sizeof t, "/api/v%u/resource/{id}/items/%u", 1 + r.below(3), r.skew(20000) % 7);
    if (r.chance(0.9)) req.str(4, t);
  }
  uint32_t status = 0;
  int32_t code = 200;
  if (msg_type != 0 || r.chance(0.2)) {
    uint32_t k = r.below(100);
    if (k < 80) { status = 0; code = (int32_t[]){200, 201, 204, 301, 302, 304}[r.below(6)]; }
    else if (k < 90) { status = 4; code = (int32_t[]){400, 401, 403, 404, 405, 429, 499}[r.below(7)]; }
    else if (k < 97) { status = 3; code = (int32_t[]){500, 502, 503, 504, 599, 600}[r.below(6)]; }
    else { status = 2; code = 0; }
    if (r.chance(0.02)) code = -32768;  // L7PROTOCOL_LOG_RESP_CODE_NONE
    if (l7proto == 100) code = r.chance(0.5) ? 0 : (int32_t)r.below(110) - 1;
    if (l7proto == 120 || l7proto == 101) code = (int32_t)r.below(200);
    resp.u(1, status);
    resp.i32(2, code);
    if (r.chance(0.1)) resp.str(3, "custom exception text");
    if (r.chance(0.1)) resp.str(4, "ok-result");
  }
  bool has_trace = !r.chance(0.10);
  if (has_trace) {
    std::string tid = hexid(r, 32);
    if (r.chance(0.7)) trace.str(1, tid);
    trace.str(2, hexid(r, 16));
    if (r.chance(0.6)) trace.str(3, hexid(r, 16));
    if (r.chance(0.3)) {
      trace.str_always(4, r.chance(0.9) ? tid : std::string());
      if (r.chance(0.3)) trace.str_always(4, hexid(r, 32));
      if (r.chance(0.1)) trace.str_always(4, hexid(r, 8));
    }
  }
  uint32_t nattr = r.below(7);
  if (r.chance(0.7)) {
    snprintf(t, sizeof t, "app-svc-%u", r.skew(30));
    ext.str(1, t);
  }
  if (r.chance(0.2)) {
    snprintf(t, sizeof t, "10.%u.%u.%u", r.below(256), r.below(256), r.below(256));
    ext.str(2, t);
  }
  if (r.chance(0.3)) ext.u(3, r.below(1u << 31));
  if (r.chance(0.3)) ext.str(4, hexid(r, 24));
  if (r.chance(0.3)) {
    snprintf(t, sizeof t, "Mozilla/5.0 (X11; Linux x86_64) agent/%u", r.skew(10));
    ext.str(6, t);
  }
  if (r.chance(0.2)) {
    snprintf(t, sizeof t, "https://svc-%u.example.com/ref/%u", r.skew(20), r.skew(5));
    ext.str(7, t);
  }
  if (r.chance(0.05)) ext.str(8, "rpc.Service/Method");
  if (r.chance(0.03)) ext.str(9, "MyBizProto");
  if (r.chance(0.2)) ext.str(10, hexid(r, 24));
  uint32_t nval = nattr;
  if ((flags & F_DUP_FIELDS) && r.chance(0.05)) nval = nattr ? nattr - 1 : 1;  // mismatched lengths
  // 60 names x 4 values each; with the bounded (vpc, pod_ns) set this keeps the tag key space small
  uint32_t name_idx[8];
  const bool hc = (flags & F_HIGH_CARD) != 0;
  for (uint32_t i = 0; i < nattr; i++) {
    name_idx[i] = hc ? r.below(8) : r.skew(40);
    snprintf(t, sizeof t, "attr_%u", name_idx[i]);
    ext.str_always(16, t);
  }
  for (uint32_t i = 0; i < nval; i++) {
    if (hc) snprintf(t, sizeof t, "user-%u", r.below(1u << 17));
    else snprintf(t, sizeof t, "val-%u", (i < nattr ? name_idx[i] : 0) * 3 + r.skew(3));
    ext.str_always(17, r.chance(0.03) ? std::string() : std::string(t));
  }
  if (r.chance(0.1)) {
    uint32_t nm = 1 + r.below(3);
    std::vector<double> vals;
    for (uint32_t i = 0; i < nm; i++) {
      snprintf(t, sizeof t, "metric_%u", r.skew(40));
      ext.str_always(18, t);
      vals.push_back((double)r.below(100000) / 7.0);
    }
    ext.packed_dbl(19, vals, !(flags & F_UNPACKED) || r.chance(0.5));
  }
  if ((flags & F_UNKNOWN_FIELDS) && r.chance(0.1)) unknown_fields(ext, r);

  out.msg(1, base);
  if (r.chance(0.9)) out.i32(9, r.chance(0.05) ? -1 : (int32_t)r.below(1 << 20));
  if (r.chance(0.9)) out.i32(10, r.chance(0.05) ? -1 : (int32_t)r.below(1 << 22));
  bool omit = (flags & F_OMIT_SUBMSG) != 0;
  if (!(omit && r.chance(0.15))) out.msg(11, req);
  if (!(omit && r.chance(0.15))) out.msg(12, resp);
  out.str(13, l7proto == 21 ? "2" : (r.chance(0.9) ? "1.1" : "1.0"));
  if (!(omit && r.chance(0.15))) out.msg(14, trace);
  if (!(omit && r.chance(0.15))) out.msg(15, ext);
  if (r.chance(0.05)) out.u(16, r.below(10000));
  out.u(17, r.below(256));
  out.u(18, (r.chance(0.05) ? 1 : 0) | (r.chance(0.03) ? 2 : 0) | (r.chance(0.03) ? 4 : 0));
  out.u(19, r.logu(20));
  out.u(20, r.logu(22));
  if (r.chance(0.03)) { out.str(21, "BIZ-0001"); out.str(22, "checkout"); out.str(23, "00"); }
  if ((flags & F_UNKNOWN_FIELDS) && r.chance(0.1)) unknown_fields(out, r);
}

// ---- metric.proto Document (message/metric.proto:14-195), as the agent's collector emits
// it: one MiniTag (single-side or edge code) + one meter.
void sparse_u(PB& pb, Rng& r, uint32_t f, int bits, double p_zero = 0.6) {
  if (!r.chance(p_zero)) pb.u(f, r.logu(bits));
}

void gen_metric(synth_ctx* c, Rng& r, uint32_t vtap, uint64_t base_ns, uint32_t flags, PB& out) {
  // Code bits (flow-metrics/tag.go:38-94)
  const uint64_t IP = 1, L3EPC = 2, MAC = 1ull << 11, GPID = 1ull << 15, PATH = 20;
  const uint64_t DIRECTION = 1ull << 40, ACLGID = 1ull << 41, PROTOCOL = 1ull << 42, SERVERPORT = 1ull << 43,
                 TAPTYPE = 1ull << 45, VTAPID = 1ull << 47, TAPSIDE = 1ull << 48, TAPPORT = 1ull << 49,
                 L7PROTOCOL = 1ull << 51, TUNNELIPID = 1ull << 62;
  uint32_t pick = r.below(100);
  // 0 network, 1 network_map, 2 application, 3 application_map, 4 traffic_policy
  int kind = pick < 25 ? 0 : pick < 55 ? 1 : pick < 72 ? 2 : pick < 94 ? 3 : 4;
  const bool edge = kind == 1 || kind == 3, app = kind == 2 || kind == 3, policy = kind == 4;
  const bool v6 = !policy && r.chance(0.10);
  Side s0, s1;
  auto mode = [&]() { uint32_t m = r.below(100); return m < 80 ? 0 : m < 92 ? 1 : 2; };
  pick_side(c, r, s0, v6, mode());
  pick_side(c, r, s1, v6, mode());

  uint64_t code;
  if (policy) code = ACLGID | TUNNELIPID | VTAPID;
  else if (edge) code = (IP << PATH) | (L3EPC << PATH) | VTAPID | TAPTYPE | TAPSIDE | PROTOCOL | SERVERPORT | TAPPORT;
  else code = IP | L3EPC | VTAPID | TAPTYPE | PROTOCOL | SERVERPORT | DIRECTION;
  if (app) code |= L7PROTOCOL;
  if (!policy && r.chance(0.2)) code |= edge ? (MAC << PATH) : MAC;   // ignored by TableID
  if (!policy && r.chance(0.5)) code |= edge ? (GPID << PATH) : GPID;  // newer agents set it themselves
  if (r.chance(0.04)) {  // codes that match no table -> DropDocCount
    switch (r.below(4)) {
      case 0: code &= ~SERVERPORT; break;
      case 1: code |= edge ? DIRECTION : TAPSIDE; break;
      case 2: code = 0; break;
      case 3: code |= ACLGID; break;
    }
  }

  PB f;  // MiniField
  uint8_t ip4be[4];
  auto put_ip = [&](uint32_t field, const Side& s) {
    if (v6) f.bytes(field, s.ip6, 16);
    else {
      ip4be[0] = (uint8_t)(s.ip4 >> 24); ip4be[1] = (uint8_t)(s.ip4 >> 16);
      ip4be[2] = (uint8_t)(s.ip4 >> 8);  ip4be[3] = (uint8_t)s.ip4;
      f.tag(field, 2); f.varint(4); f.b.insert(f.b.end(), ip4be, ip4be + 4);  // always 4 bytes, even 0.0.0.0
    }
  };
  put_ip(1, s0);
  if (edge || r.chance(0.1)) put_ip(2, s1);
  f.u(3, r.below(policy ? 4 : 40));
  f.u(4, v6 ? 1 : 0);
  int32_t epc0 = s0.epc, epc1 = s1.epc;
  if (r.chance(0.01)) epc0 = -(int32_t)(3 + r.below(40000));
  f.i32(5, epc0);
  if (edge) f.i32(6, epc1);
  if (r.chance(0.5)) f.u(7, s0.mac);
  if (edge && r.chance(0.5)) f.u(8, s1.mac);
  static const uint32_t kDirs[] = {1, 2, 4, 0, 1 | 8, 2 | 8, 1 | 16, 2 | 16, 1 | 40, 2 | 40, 1 | 48, 2 | 48, 3, 7};
  uint32_t direction = kDirs[r.skew(14)];
  f.u(9, direction);
  uint32_t tap_side = kTapSides[r.below(17)];
  if (r.chance(0.3)) tap_side = r.chance(0.5) ? 1 : 2;  // c / s: the region filter applies to these
  f.u(10, tap_side);
  f.u(11, r.chance(0.7) ? 6 : r.chance(0.7) ? 17 : r.below(256));
  if (policy || r.chance(0.02)) f.u(12, 1 + r.below(3000));
  uint32_t port = r.chance(0.85) ? c->svc_ports[r.below((uint32_t)c->svc_ports.size())] : r.below(65536);
  if (r.chance(0.1)) port = 0;
  f.u(13, (flags & F_WIDE_VARINT) && r.chance(0.05) ? (uint64_t)port | 0x70000ull : port);
  f.u(14, vtap);
  uint64_t tap_port = (uint64_t)(r.next() & 0xffffffffu) | (uint64_t)r.below(16) << 32 | (uint64_t)r.below(4) << 36 |
                      (uint64_t)r.below(8) << 60;
  if (edge || r.chance(0.2)) f.u(15, tap_port);
  f.u(16, r.chance(0.9) ? 3 : 1 + r.below(255));
  if (app) f.u(17, r.chance(0.6) ? 20 : r.chance(0.5) ? 120 : r.below(130));
  if (r.chance(0.6)) f.u(20, s0.gpid);
  if (edge && r.chance(0.6)) f.u(21, s1.gpid);
  uint32_t sig = r.below(100);
  f.u(22, sig < 55 ? 0 : sig < 85 ? 3 : sig < 95 ? 4 : r.below(8));
  if (app) {
    char tmp[48];
    if (r.chance(0.7)) { snprintf(tmp, sizeof tmp, "svc-%u", r.skew(60)); f.str(23, tmp); }
    if (r.chance(0.3)) { snprintf(tmp, sizeof tmp, "inst-%u", r.skew(25)); f.str(24, tmp); }
    if (r.chance(0.8)) { snprintf(tmp, sizeof tmp, "/api/v%u/resource/%u", 1 + r.below(3), r.skew(300)); f.str(25, tmp); }
    if (r.chance(0.1)) f.u(28, 1 + r.below(5));
  }
  if (r.chance(0.3)) {
    const Side& ps = (edge && (direction & 7) == 2) ? s1 : s0;
    f.u(27, ps.pod_id);
  }
  if ((flags & F_UNKNOWN_FIELDS) && r.chance(0.1)) unknown_fields(f, r);

  PB tag;
  tag.msg(1, f);
  tag.u(2, code);
  if ((flags & F_DUP_FIELDS) && r.chance(0.1)) {  // a second MiniField occurrence merges into the first
    PB f2;
    f2.u(14, vtap);
    f2.u(11, 6);
    tag.msg(1, f2);
  }

  PB meter;
  const bool omit = (flags & F_OMIT_SUBMSG) != 0;
  uint32_t meter_id = policy ? 4 : app ? 5 : 1;
  if (r.chance(0.01)) meter_id = policy ? 1 : 4;  // meter / table mismatch -> dropped
  if ((flags & F_CORRUPT) && r.chance(0.0005)) meter_id = r.chance(0.5) ? 2 : 7;  // unknown meter: frame abandoned
  meter.u(1, (flags & F_WIDE_VARINT) && r.chance(0.05) ? (uint64_t)meter_id + 256 : meter_id);  // uint8(MeterId)
  if (meter_id == 1 || r.chance(0.02)) {
    PB fm, tr, la, pe, an, fl;
    for (uint32_t i = 1; i <= 10; i++) sparse_u(tr, r, i, r.chance(0.02) ? 62 : 40);
    for (uint32_t i = 11; i <= 14; i++) sparse_u(tr, r, i, 24);
    if ((flags & F_WIDE_VARINT) && r.chance(0.1)) tr.u(11, 0x100000000ull + r.below(1000));  // uint32 keeps the low bits
    sparse_u(tr, r, 15, 8);
    for (uint32_t i = 1; i <= 6; i++) sparse_u(la, r, i, 30);
    sparse_u(la, r, 19, 30);
    for (uint32_t i = 7; i <= 12; i++) sparse_u(la, r, i, r.chance(0.05) ? 63 : 44);  // > 2^53: f64 rounding
    sparse_u(la, r, 20, 44);
    for (uint32_t i = 13; i <= 18; i++) sparse_u(la, r, i, 20);
    sparse_u(la, r, 21, 20);
    for (uint32_t i = 1; i <= 4; i++) sparse_u(pe, r, i, 30);
    sparse_u(pe, r, 5, 16);
    sparse_u(pe, r, 6, 16);
    for (uint32_t i = 1; i <= 12; i++) sparse_u(an, r, i, r.chance(0.01) ? 64 : 24, 0.75);
    for (uint32_t i = 13; i <= 15; i++) sparse_u(an, r, i, 32, 0.75);
    sparse_u(an, r, 20, 20, 0.8);
    sparse_u(an, r, 21, 20, 0.8);
    sparse_u(fl, r, 1, 16, 0.5);
    if (!(omit && r.chance(0.2))) fm.msg(1, tr);
    if (!(omit && r.chance(0.2))) fm.msg(2, la);
    if (!(omit && r.chance(0.2))) fm.msg(3, pe);
    if (!(omit && r.chance(0.2))) fm.msg(4, an);
    if (!(omit && r.chance(0.2))) fm.msg(5, fl);
    if ((flags & F_DUP_FIELDS) && r.chance(0.1)) {  // split sub-message: merged, last scalar wins
      PB tr2;
      tr2.u(1, 77);
      tr2.u(15, 200);
      fm.msg(1, tr2);
    }
    meter.msg(2, fm);
  }
  if (meter_id == 4 || r.chance(0.02)) {
    PB um;
    for (uint32_t i = 1; i <= 8; i++) sparse_u(um, r, i, 40, 0.3);
    meter.msg(3, um);
  }
  if (meter_id == 5 || r.chance(0.02)) {
    PB am, t, l, a;
    sparse_u(t, r, 1, 24, 0.2);
    sparse_u(t, r, 2, 24, 0.2);
    sparse_u(t, r, 3, 8, 0.3);
    sparse_u(l, r, 1, 30, 0.3);
    sparse_u(l, r, 2, r.chance(0.05) ? 63 : 44, 0.3);
    sparse_u(l, r, 3, 20, 0.3);
    sparse_u(a, r, 1, r.chance(0.02) ? 32 : 16, 0.7);
    sparse_u(a, r, 2, r.chance(0.02) ? 32 : 16, 0.7);
    sparse_u(a, r, 3, 16, 0.7);
    if (!(omit && r.chance(0.2))) am.msg(1, t);
    if (!(omit && r.chance(0.2))) am.msg(2, l);
    if (!(omit && r.chance(0.2))) am.msg(3, a);
    meter.msg(4, am);
  }
  if ((flags & F_UNKNOWN_FIELDS) && r.chance(0.05)) unknown_fields(meter, r);

  // 60 consecutive seconds; a few documents an hour later (second time bucket / lru expiry)
  uint32_t ts = (uint32_t)(base_ns / 1000000000ull) + r.below(60) + (r.chance(0.01) ? 3600 : 0);
  out.u(1, ts);
  out.msg(2, tag);
  out.msg(3, meter);
  // FLAG_PER_SECOND_METRICS | an unused bit; traffic_policy has no per-second table (a
  // per-second policy document would index past metricsTableCodes in the reference)
  out.u(4, (!policy && r.chance(0.5) ? 1 : 0) | (r.chance(0.02) ? 2 : 0));
  if ((flags & F_DUP_FIELDS) && r.chance(0.05)) out.u_always(1, ts);
  if ((flags & F_UNKNOWN_FIELDS) && r.chance(0.1)) unknown_fields(out, r);
}

}  // namespace

extern "C" {

synth_ctx* synth_new(uint64_t seed, uint32_t n_pods) {
  auto* c = new synth_ctx();
  c->seed = seed;
  c->P = std::max<uint32_t>(64, n_pods);
  build_dict(c);
  return c;
}
void synth_free(synth_ctx* c) { delete c; }
const dfgpu_dict_snapshot* synth_dict(synth_ctx* c) { return &c->snap; }
int synth_geo(synth_ctx* c, const dfgpu_geo_entry** e, uint64_t* n, const char* const** names, uint32_t* n_names) {
  *e = c->geo.data();
  *n = c->geo.size();
  *names = c->geo_name_ptrs.data();
  *n_names = (uint32_t)c->geo_name_ptrs.size();
  return 0;
}

// vtap_mod/vtap_rem: only agents with vtap_id % vtap_mod == vtap_rem send frames (the
// per-GPU shard of a vtap-sharded firehose); vtap_mod <= 1 = every agent.
synth_batch* synth_gen(synth_ctx* c, int msg_type, uint64_t seed, uint64_t nrecords, uint32_t flags,
                       uint32_t max_frame_bytes, uint16_t org_id, uint32_t vtap_mod, uint32_t vtap_rem) {
  auto* b = new synth_batch();
  Rng r(seed * 1000003ull + (uint64_t)msg_type);
  if (!max_frame_bytes) max_frame_bytes = 256 * 1024;
  uint64_t base_ns = 1790000000ull * 1000000000ull;
  uint64_t made = 0;
  b->data.reserve((size_t)(nrecords * (msg_type == DFGPU_MSG_PROTOCOLLOG ? 460 : 330)));
  PB pb;
  while (made < nrecords) {
    dfgpu_frame fr;
    memset(&fr, 0, sizeof fr);
    fr.offset = b->data.size();
    fr.vtap_id = (uint16_t)(1 + r.below(c->n_vtap));
    if (vtap_mod > 1) {
      uint32_t first = vtap_rem % vtap_mod ? vtap_rem % vtap_mod : vtap_mod;
      uint32_t count = first <= c->n_vtap ? (c->n_vtap - first) / vtap_mod + 1 : 1;
      fr.vtap_id = (uint16_t)(first + vtap_mod * r.below(count));
    }
    fr.org_id = org_id;
    fr.team_id = 1 + (fr.vtap_id % 3);
    // agents flush by size or time: vary the frame fill
    uint32_t budget = r.chance(0.8) ? max_frame_bytes : 512 + r.below(max_frame_bytes);
    uint32_t used = 0;
    while (made < nrecords) {
      pb.b.clear();
      if (msg_type == DFGPU_MSG_TAGGEDFLOW) gen_l4(c, r, fr.vtap_id, base_ns, flags, pb);
      else if (msg_type == DFGPU_MSG_PROTOCOLLOG) gen_l7(c, r, fr.vtap_id, base_ns, flags, pb);
      else if (msg_type == DFGPU_MSG_METRICS) gen_metric(c, r, fr.vtap_id, base_ns, flags, pb);
      else { delete b; return nullptr; }
      uint32_t n = (uint32_t)pb.b.size();
      if (used && used + 4 + n > budget) break;
      uint8_t le[4] = {(uint8_t)n, (uint8_t)(n >> 8), (uint8_t)(n >> 16), (uint8_t)(n >> 24)};
      b->data.insert(b->data.end(), le, le + 4);
      b->data.insert(b->data.end(), pb.b.begin(), pb.b.end());
      used += 4 + n;
      made++;
    }
    fr.len = used;
    if ((flags & F_CORRUPT) && r.chance(0.05) && used > 64) {
      uint32_t kind = r.below(4);
      uint8_t* p = b->data.data() + fr.offset;
      if (kind == 0) fr.len = used - 1 - r.below(40);                 // truncated frame
      else if (kind == 1) p[used / 2] ^= 0x80 | (uint8_t)r.below(128); // damaged byte mid-frame
      else if (kind == 2) { p[0] = 0xff; p[1] = 0xff; p[2] = 0xff; p[3] = 0x7f; }  // absurd length
      else { uint32_t o = used / 3; p[o] = 0xff; p[o + 1] = 0xff; p[o + 2] = 0xff; p[o + 3] = 0xff; p[o + 4] = 0xff; }
    }
    b->frames.push_back(fr);
    if ((flags & F_CORRUPT) && r.chance(0.02)) {  // empty frame
      dfgpu_frame e = fr;
      e.offset = b->data.size();
      e.len = 0;
      b->frames.push_back(e);
    }
  }
  b->nrecords = made;
  return b;
}

void synth_batch_view(const synth_batch* b, synth_frames* out) {
  out->data = b->data.data();
  out->nbytes = b->data.size();
  out->frames = b->frames.data();
  out->nframes = (uint32_t)b->frames.size();
  out->pad = 0;
  out->nrecords = b->nrecords;
}
void synth_batch_free(synth_batch* b) { delete b; }

}  // extern "C"
