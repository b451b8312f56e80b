// Library plumbing behind include/omc.h: error text, device init, CUDA-graph sweep runner, sample store.
#include "../../include/omc.h"
#include "omc_common.cuh"
#include "omc_internal.h"
#include <stdarg.h>
#include <vector>

static thread_local char g_err[1024] = "";
static int g_sm_count = 0;

void omc_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
int omc_sm_count() {
  if (g_sm_count == 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess ||
        cudaDeviceGetAttribute(&g_sm_count, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess)
      g_sm_count = 148;  // B200
  }
  return g_sm_count;
}

struct omc_graph {
  cudaGraph_t graph = nullptr;
  cudaGraphExec_t exec = nullptr;
};

namespace {
__global__ void counter_add_kernel(unsigned long long* c, unsigned long long inc) { *c += inc; }
__global__ void counter_add2_kernel(unsigned long long* c0, unsigned long long inc0, unsigned long long* c1,
                                    unsigned long long inc1) {
  *c0 += inc0;
  *c1 += inc1;
}

// RING = false: row `it` of a [max_iter, count] store (rows beyond it are dropped); RING = true: slot it % max_iter of a
// ring of max_iter slabs that a copy stream drains to the host while the next sweeps run
template <bool RING>
__global__ void store_copy_kernel(const double* __restrict__ src, double* __restrict__ dst, long long count,
                                  const unsigned long long* iter, long long max_iter) {
  unsigned long long it = *iter;
  if (RING) it %= (unsigned long long)max_iter;
  else if ((long long)it >= max_iter) return;
  double* d = dst + it * count;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (long long)gridDim.x * blockDim.x)
    d[i] = src[i];
}
}  // namespace

extern "C" {

int omc_abi_version(void) { return OMC_ABI_VERSION; }
const char* omc_last_error(void) { return g_err; }

int omc_device_init(int device) {
  int n = 0;
  OMC_CHECK_CUDA(cudaGetDeviceCount(&n));
  OMC_REQUIRE(device >= 0 && device < n, "omc_device_init: device %d not in [0,%d)", device, n);
  OMC_CHECK_CUDA(cudaSetDevice(device));
  cudaDeviceProp prop;
  OMC_CHECK_CUDA(cudaGetDeviceProperties(&prop, device));
  OMC_REQUIRE(prop.major == 10, "omc_device_init: libomc is built for sm_100a only, device is sm_%d%d", prop.major,
              prop.minor);
  g_sm_count = prop.multiProcessorCount;
  return 0;
}
int omc_device_sm_count(void) { return omc_sm_count(); }

int omc_counter_add(unsigned long long* counter, unsigned long long inc, void* stream) {
  OMC_REQUIRE(counter, "omc_counter_add: null counter");
  counter_add_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(counter, inc);
  OMC_LAUNCH_CHECK();
  return 0;
}

int omc_counter_add2(unsigned long long* c0, unsigned long long inc0, unsigned long long* c1, unsigned long long inc1,
                     void* stream) {
  OMC_REQUIRE(c0 && c1, "omc_counter_add2: null counter");
  counter_add2_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(c0, inc0, c1, inc1);
  OMC_LAUNCH_CHECK();
  return 0;
}

int omc_graph_capture_begin(void* stream) {
  OMC_CHECK_CUDA(cudaStreamBeginCapture((cudaStream_t)stream, cudaStreamCaptureModeThreadLocal));
  return 0;
}
int omc_graph_capture_end(void* stream, omc_graph_t** out) {
  OMC_REQUIRE(out, "omc_graph_capture_end: null out");
  omc_graph* g = new omc_graph();
  cudaError_t e = cudaStreamEndCapture((cudaStream_t)stream, &g->graph);
  if (e != cudaSuccess) {
    delete g;
    omc_set_error("cudaStreamEndCapture: %s", cudaGetErrorString(e));
    return (int)e;
  }
  e = cudaGraphInstantiate(&g->exec, g->graph, 0);
  if (e != cudaSuccess) {
    cudaGraphDestroy(g->graph);
    delete g;
    omc_set_error("cudaGraphInstantiate: %s", cudaGetErrorString(e));
    return (int)e;
  }
  *out = g;
  return 0;
}
int omc_graph_launch(omc_graph_t* g, void* stream, long long times) {
  OMC_REQUIRE(g && g->exec, "omc_graph_launch: null graph");
  for (long long i = 0; i < times; ++i) OMC_CHECK_CUDA(cudaGraphLaunch(g->exec, (cudaStream_t)stream));
  return 0;
}
int omc_graph_destroy(omc_graph_t* g) {
  if (!g) return 0;
  if (g->exec) cudaGraphExecDestroy(g->exec);
  if (g->graph) cudaGraphDestroy(g->graph);
  delete g;
  return 0;
}
int omc_graph_num_kernel_nodes(omc_graph_t* g, long long* out) {
  OMC_REQUIRE(g && g->graph && out, "omc_graph_num_kernel_nodes: null argument");
  size_t n = 0;
  OMC_CHECK_CUDA(cudaGraphGetNodes(g->graph, nullptr, &n));
  std::vector<cudaGraphNode_t> nodes(n);
  if (n) OMC_CHECK_CUDA(cudaGraphGetNodes(g->graph, nodes.data(), &n));
  long long k = 0;
  for (size_t i = 0; i < n; ++i) {
    cudaGraphNodeType t;
    OMC_CHECK_CUDA(cudaGraphNodeGetType(nodes[i], &t));
    if (t == cudaGraphNodeTypeKernel) ++k;
  }
  *out = k;
  return 0;
}

// ref: mcmc.py:97-111 — iterations numbered -n_burn..n_iter-1, n_thin sweeps each, store after non-burn iterations
int omc_run_schedule(omc_graph_t* sweep, omc_graph_t* store, void* stream, long long n_burn, long long n_iter,
                     long long n_thin) {
  OMC_REQUIRE(sweep && sweep->exec, "omc_run_schedule: null sweep graph");
  OMC_REQUIRE(n_burn >= 0 && n_iter >= 0 && n_thin >= 1, "omc_run_schedule: bad schedule %lld/%lld/%lld", n_burn,
              n_iter, n_thin);
  cudaStream_t st = (cudaStream_t)stream;
  for (long long it = -n_burn; it < n_iter; ++it) {
    for (long long t = 0; t < n_thin; ++t) OMC_CHECK_CUDA(cudaGraphLaunch(sweep->exec, st));
    if (it >= 0 && store && store->exec) OMC_CHECK_CUDA(cudaGraphLaunch(store->exec, st));
  }
  return 0;
}

int omc_store_copy(const double* src, double* dst, long long count, const unsigned long long* iter_counter,
                   long long max_iter, void* stream) {
  OMC_REQUIRE(src && dst && iter_counter && count >= 0, "omc_store_copy: bad argument");
  if (count == 0) return 0;
  long long blocks = (count + 255) / 256;
  if (blocks > 148 * 8) blocks = 148 * 8;
  store_copy_kernel<false><<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(src, dst, count, iter_counter, max_iter);
  OMC_LAUNCH_CHECK();
  return 0;
}

int omc_store_copy_ring(const double* src, double* dst, long long count, const unsigned long long* iter_counter,
                        long long ring, void* stream) {
  OMC_REQUIRE(src && dst && iter_counter && count >= 0 && ring >= 1, "omc_store_copy_ring: bad argument");
  if (count == 0) return 0;
  long long blocks = (count + 255) / 256;
  if (blocks > 148 * 8) blocks = 148 * 8;
  store_copy_kernel<true><<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(src, dst, count, iter_counter, ring);
  OMC_LAUNCH_CHECK();
  return 0;
}

}  // extern "C"
