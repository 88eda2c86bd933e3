/* l7.cu — placeholder until the L7 kernels land (same commit series). */
#include "pipeline.cuh"
struct dfgpu_ctx;
namespace dfgpu {
void launch_l7_pass(const BatchParams&, int, cudaStream_t) {}
int l7_flow_tags(dfgpu_ctx*, cudaStream_t) { return 0; }
int l7_fetch_tables(dfgpu_ctx*, dfgpu_batch*) { return 0; }
void l7_streams(BatchParams& bp) { bp.nslots = 1; bp.nstreams = 0; bp.n_lc = 0; }
int l7_lc_space(int) { return 0; }
const std::vector<std::string>& l7_lc_names(dfgpu_ctx*, int) { static std::vector<std::string> v; return v; }
}
extern "C" void dfgpu_l7_free_state(dfgpu_ctx*) {}
