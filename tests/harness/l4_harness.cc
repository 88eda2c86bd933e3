// tests/harness/l4_harness.cc — TEST INFRASTRUCTURE, not part of libdfgpu.so (see l7_harness.cc).
// The L4 decode core of the product (pb_tok.cuh tokenizer + l4_tok.h table + l4_row.cuh row builder: the exact
// source the CUDA kernel l4.cu instantiates on the device) compiled for the host.
#include <cstring>
#include <vector>

#include "../../deepflow_b200/csrc/l4_row.cuh"

using namespace dfgpu;

namespace {
struct HostSink {
  uint32_t w[W4_COUNT];
  uint64_t elems[L4E_COUNT];
  void u32(uint32_t s, uint32_t v) { w[s] = v; }
  void u64(uint32_t s, uint32_t lo, uint32_t hi) { w[s] = lo; w[s + 1] = hi; }
  void add(uint32_t s, uint32_t v) { w[s] += v; }
  uint32_t word(uint32_t s) const { return w[s]; }
  void ref(uint32_t, uint32_t, uint32_t) {}
  void elem(uint32_t, uint32_t, uint32_t, uint32_t) {}
  void velem(uint32_t base, uint32_t i, uint32_t lo, uint32_t hi) { if (i < L4E_CAP) elems[base + i] = (uint64_t)hi << 32 | lo; }
};
struct HostView {
  const HostSink* s;
  uint32_t w(int i) const { return s->w[i]; }
  uint64_t w64(int i) const { return (uint64_t)s->w[i + 1] << 32 | s->w[i]; }
};
struct HostEmit {
  uint8_t** fixp; const uint8_t* in; uint32_t* kgp; uint32_t rec;
  template <class T> void fix(int col, T v) { if (fixp[col]) memcpy(fixp[col] + (size_t)rec * sizeof(T), &v, sizeof(T)); }
  void ip6(int col, uint32_t off, uint32_t len) {
    uint8_t b[16] = {0};
    for (uint32_t i = 0; i < len && i < 16; i++) b[i] = in[off + i];
    if (fixp[col]) memcpy(fixp[col] + (size_t)rec * 16, b, 16);
  }
  void kg(uint32_t vtap, uint32_t tap_side, uint32_t proto, uint32_t port_dst, bool v6, bool vip0, bool vip1, uint32_t ip0,
          uint32_t ip1, uint32_t epc0, uint32_t epc1, uint64_t mac0, uint64_t mac1, uint32_t gpid0, uint32_t gpid1, uint32_t,
          uint32_t, uint32_t, uint32_t) {
    uint32_t* k = kgp + (size_t)rec * 16;
    k[0] = vtap; k[1] = tap_side; k[2] = proto; k[3] = port_dst; k[4] = v6 | vip0 << 1 | vip1 << 2; k[5] = ip0; k[6] = ip1;
    k[7] = epc0; k[8] = epc1; k[9] = (uint32_t)mac0; k[10] = (uint32_t)mac1; k[11] = gpid0; k[12] = gpid1;
  }
};
std::vector<uint32_t> g_tbl;
}  // namespace

extern "C" {
int harness_l4_ncols() { return L4C_COUNT; }
// info [8][nrec]: tside, ip_src, ip_dst, dom_off, dom_len, n_acl, n_agg, agg_chars; elems [nrec][16] u64
int harness_l4(const uint8_t* in, uint64_t nbytes, const uint32_t* rec_off, const uint32_t* rec_len, uint32_t nrec,
               uint32_t analyzer_id, uint32_t id_counter, uint8_t* ok, uint8_t** fix, uint32_t* info, uint64_t* elems,
               uint32_t* kg) {
  if (g_tbl.empty()) g_tbl = l4_tok_table();
  TokWin win;
  win.init(in, nbytes);
  for (uint32_t r = 0; r < nrec; r++) {
    HostSink sink;
    memset(&sink, 0, sizeof sink);
    const bool good = tk_decode<L4_MAXF>(win, rec_off[r], rec_len[r], g_tbl.data(), sink, 0u);
    ok[r] = good ? 1 : 0;
    if (!good) memset(&sink, 0, sizeof sink);
    HostView view{&sink};
    HostEmit em{fix, in, kg, r};
    const L4RowInfo ri = l4_build_row(view, em, analyzer_id, id_counter + r + 1);
    const uint32_t v[8] = {ri.tside, ri.ip_src, ri.ip_dst, ri.dom_off, ri.dom_len, ri.n_acl, ri.n_agg, ri.agg_chars};
    for (int k = 0; k < 8; k++) info[(size_t)k * nrec + r] = v[k];
    memcpy(elems + (size_t)r * L4E_COUNT, sink.elems, sizeof sink.elems);
  }
  return 0;
}
}  // extern "C"
