// tests/harness/l7_harness.cc — TEST INFRASTRUCTURE, not part of libdfgpu.so.
// Compiles the L7 decode core of the product (pb_tok.cuh tokenizer + l7_row.cuh row builder: the exact
// source the CUDA kernel l7_pass.cu instantiates on the device) for the HOST with plain g++, so that the
// per-record logic can be checked against the oracle in the `-m "not gpu"` tier.  The product library never
// runs this code on the CPU: without a CUDA device dfgpu_init fails.
#include <cstring>
#include <string>
#include <vector>

#include "../../deepflow_b200/csrc/l7_tables_host.h"

using namespace dfgpu;

namespace {
struct HostSink {
  uint32_t w[W7_COUNT];
  uint64_t refs[R7_COUNT];
  uint32_t mask;
  void u32(uint32_t s, uint32_t v) { w[s] = v; }
  void u64(uint32_t s, uint32_t lo, uint32_t hi) { w[s] = lo; w[s + 1] = hi; }
  void add(uint32_t s, uint32_t v) { w[s] += v; }
  uint32_t word(uint32_t s) const { return w[s]; }
  void ref(uint32_t r, uint32_t off, uint32_t len) { refs[r] = (uint64_t)len << 32 | off; mask |= 1u << r; }
  void elem(uint32_t, uint32_t, uint32_t, uint32_t) {}
  void velem(uint32_t, uint32_t, uint32_t, uint32_t) {}
};
struct HostView {
  const HostSink* s;
  uint32_t w(int i) const { return s->w[i]; }
  uint64_t w64(int i) const { return (uint64_t)s->w[i + 1] << 32 | s->w[i]; }
  SRef32 ref(int r) const {
    const uint64_t v = ((s->mask >> r) & 1u) ? s->refs[r] : 0;
    return SRef32{(uint32_t)v, (uint32_t)(v >> 32)};
  }
};
struct HostEmit {
  uint8_t** fixp; uint8_t** nullp; const uint8_t* in;
  uint32_t* str_off; uint32_t* str_len; uint32_t* aux_off; uint32_t* aux_len; uint32_t* kgp;
  uint32_t nrec, rec, meta;
  template <class T> void fix_(int col, T v) { if (fixp[col]) memcpy(fixp[col] + (size_t)rec * sizeof(T), &v, sizeof(T)); }
  template <class T> void fix(int col, T v) { fix_<T>(col, v); }
  void ip6(int col, SRef32 r) {
    uint8_t b[16] = {0};
    for (uint32_t i = 0; i < r.len && i < 16; i++) b[i] = in[r.off + i];
    if (fixp[col]) memcpy(fixp[col] + (size_t)rec * 16, b, 16);
  }
  template <class T> void nullable(int col, bool has, T v) {
    fix_<T>(col, has ? v : T(0));
    if (nullp[col]) nullp[col][rec] = has ? 0 : 1;
  }
  void str(int i, SRef32 r) { str_off[(size_t)i * nrec + rec] = r.off; str_len[(size_t)i * nrec + rec] = r.len; }
  void aux(int k, SRef32 r) { aux_off[(size_t)k * nrec + rec] = r.off; aux_len[(size_t)k * nrec + rec] = r.len; }
  void flags(bool has_ext, bool slow) { meta = (has_ext ? 2u : 0u) | (slow ? 4u : 0u); }
  void kg(uint32_t vtap, uint32_t tap_side, uint32_t proto, uint32_t port_dst, bool v6, bool vip0, bool vip1, uint32_t ip0,
          uint32_t ip1, uint32_t epc0, uint32_t epc1, uint64_t mac0, uint64_t mac1, uint32_t gpid0, uint32_t gpid1,
          uint32_t pod0, uint32_t pod1, uint32_t pid0, uint32_t pid1, SRef32, SRef32) {
    uint32_t* k = kgp + (size_t)rec * 16;
    k[0] = vtap; k[1] = tap_side; k[2] = proto; k[3] = port_dst; k[4] = v6 | vip0 << 1 | vip1 << 2; k[5] = ip0; k[6] = ip1;
    k[7] = epc0; k[8] = epc1; k[9] = (uint32_t)mac0; k[10] = (uint32_t)mac1; k[11] = gpid0 ^ gpid1; k[12] = pod0; k[13] = pod1;
    k[14] = pid0; k[15] = pid1;
  }
};
std::vector<std::string> g_cs;
std::vector<uint8_t> g_blob;
std::vector<uint32_t> g_off;
L7Tables g_tb;
std::vector<uint32_t> g_tbl;
void init_tables() {
  if (!g_tbl.empty()) return;
  memset(&g_tb, 0, sizeof g_tb);
  l7_build_tables(g_cs, g_tb);
  l7_layout_strings(g_cs, g_blob, g_off);
  g_tb.cs_blob = g_blob.data();
  g_tb.cs_off = g_off.data();
  g_tbl = l7_tok_table();
}
}  // namespace

extern "C" {

// the constant-string blob the REF_CS references point into
const uint8_t* harness_l7_cs_blob(uint64_t* n) {
  init_tables();
  *n = g_blob.size();
  return g_blob.data();
}
int harness_l7_ncols() { return L7C_COUNT; }

// Decodes nrec records of `in` (8-byte aligned, padded by >= 16 bytes) and builds their rows.
//   ok[nrec]; fix / nulls: L7C_COUNT pointers (NULL = not wanted), row = record index;
//   str_off / str_len [22][nrec] (REF_CS bit: constant blob); aux_* [4][nrec]; arr [7][nrec];
//   info [4][nrec]: tside, l7_protocol, trace_len, meta;  thash [nrec];  kg [nrec][16]
int harness_l7(const uint8_t* in, uint64_t nbytes, const uint32_t* rec_off, const uint32_t* rec_len, uint32_t nrec,
               uint32_t nat, uint32_t analyzer_id, uint32_t id_counter, uint8_t* ok, uint8_t** fix, uint8_t** nulls,
               uint32_t* str_off, uint32_t* str_len, uint32_t* aux_off, uint32_t* aux_len, uint32_t* arr, uint32_t* info,
               uint64_t* thash, uint32_t* kg) {
  init_tables();
  TokWin win;
  win.init(in, nbytes);
  for (uint32_t r = 0; r < nrec; r++) {
    HostSink sink;
    memset(&sink, 0, sizeof sink);
    const bool good = tk_decode<L7_MAXF>(win, rec_off[r], rec_len[r], g_tbl.data(), sink, nat ? ~0u : 0u);
    ok[r] = good ? 1 : 0;
    if (!good) memset(&sink, 0, sizeof sink);
    HostView view{&sink};
    HostEmit em{fix, nulls, in, str_off, str_len, aux_off, aux_len, kg, nrec, r, 0};
    const L7RowInfo ri = l7_build_row(g_tb, in, view, em, nat, analyzer_id, id_counter + r + 1);
    uint32_t* a = arr + r;
    a[0 * (size_t)nrec] = ri.z_an_e; a[1 * (size_t)nrec] = ri.z_an_b; a[2 * (size_t)nrec] = ri.z_av_e;
    a[3 * (size_t)nrec] = ri.z_av_b; a[4 * (size_t)nrec] = ri.z_mn_e; a[5 * (size_t)nrec] = ri.z_mn_b;
    a[6 * (size_t)nrec] = ri.z_mv_e;
    info[0 * (size_t)nrec + r] = ri.tside;
    info[1 * (size_t)nrec + r] = ri.l7_protocol;
    info[2 * (size_t)nrec + r] = ri.trace_len;
    info[3 * (size_t)nrec + r] = em.meta;
    thash[r] = ri.trace_hash;
  }
  return 0;
}

}  // extern "C"
