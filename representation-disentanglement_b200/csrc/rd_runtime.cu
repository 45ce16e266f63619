// rd_runtime.cu — the two runtime services around the kernels that SURVEY §8(b) lists in the C ABI, for callers that are not Python:
//   rd_graph_*: capture a sequence of rd_* launches on a stream into a CUDA graph and replay it (what rd_b200.trainer does with
//               torch.cuda.CUDAGraph: the whole training iteration is one graph launch);
//   rd_ddp_*  : data-parallel gradient averaging — one communicator per process / GPU over NCCL (NVLink 5 / NVSwitch), in-place
//               bucket all-reduce of the flat fp32 gradient buffer (ncclAvg), broadcast of rank 0's state.  libnccl.so.2 is resolved at
//               run time (dlopen; RD_B200_NCCL_LIB overrides the name) so the library links without NCCL headers and picks up the
//               same NCCL a hosting PyTorch process has already loaded.
#include <dlfcn.h>
#include "rd_common.cuh"

struct rd_graph {
  cudaGraph_t graph;
  cudaGraphExec_t exec;
};

extern "C" int rd_graph_begin(rd_ctx* ctx, rd_stream st) {
  RD_CUDA(ctx, cudaStreamBeginCapture((cudaStream_t)st, cudaStreamCaptureModeThreadLocal));
  return RD_OK;
}
extern "C" int rd_graph_end(rd_ctx* ctx, rd_stream st, rd_graph** out) {
  if (!out) RD_FAIL(ctx, RD_ERR_ARG, "graph_end: null out");
  cudaGraph_t g = nullptr;
  RD_CUDA(ctx, cudaStreamEndCapture((cudaStream_t)st, &g));
  cudaGraphExec_t e = nullptr;
  cudaError_t err = cudaGraphInstantiate(&e, g, 0);
  if (err != cudaSuccess) {
    cudaGraphDestroy(g);
    RD_FAIL(ctx, RD_ERR_CUDA, "cudaGraphInstantiate failed: %s", cudaGetErrorString(err));
  }
  rd_graph* r = new rd_graph();
  r->graph = g; r->exec = e;
  *out = r;
  return RD_OK;
}
extern "C" int rd_graph_launch(rd_ctx* ctx, rd_graph* g, rd_stream st) {
  if (!g) RD_FAIL(ctx, RD_ERR_ARG, "graph_launch: null graph");
  RD_CUDA(ctx, cudaGraphLaunch(g->exec, (cudaStream_t)st));
  return RD_OK;
}
extern "C" int rd_graph_node_count(rd_ctx* ctx, rd_graph* g, int64_t* kernels, int64_t* total) {
  if (!g) RD_FAIL(ctx, RD_ERR_ARG, "graph_node_count: null graph");
  size_t n = 0;
  RD_CUDA(ctx, cudaGraphGetNodes(g->graph, nullptr, &n));
  cudaGraphNode_t* nodes = n ? new cudaGraphNode_t[n] : nullptr;
  int64_t k = 0;
  if (n) {
    cudaError_t err = cudaGraphGetNodes(g->graph, nodes, &n);
    if (err != cudaSuccess) { delete[] nodes; RD_FAIL(ctx, RD_ERR_CUDA, "cudaGraphGetNodes failed: %s", cudaGetErrorString(err)); }
    for (size_t i = 0; i < n; ++i) {
      cudaGraphNodeType t;
      if (cudaGraphNodeGetType(nodes[i], &t) == cudaSuccess && t == cudaGraphNodeTypeKernel) ++k;
    }
    delete[] nodes;
  }
  if (kernels) *kernels = k;
  if (total) *total = (int64_t)n;
  return RD_OK;
}
extern "C" int rd_graph_destroy(rd_ctx* ctx, rd_graph* g) {
  if (!g) return RD_OK;
  cudaGraphExecDestroy(g->exec);
  cudaGraphDestroy(g->graph);
  delete g;
  return RD_OK;
}

// Zero-fill as a memset node (cudaMemsetAsync): no kernel launch, graph capturable — the workspaces the backward kernels accumulate
// into (bias-gradient rows, zero-padded transposed weights) are cleared with it.
extern "C" int rd_zero(rd_ctx* ctx, void* p, int64_t bytes, rd_stream st) {
  if (bytes > 0) RD_CUDA(ctx, cudaMemsetAsync(p, 0, (size_t)bytes, (cudaStream_t)st));
  return RD_OK;
}

// ------------------------------------------------------------------------------------------------ NCCL, resolved at run time
namespace {
typedef struct { char internal[128]; } nccl_uid;
typedef void* nccl_comm;
typedef int (*fn_get_uid)(nccl_uid*);
typedef int (*fn_init_rank)(nccl_comm*, int, nccl_uid, int);
typedef int (*fn_allreduce)(const void*, void*, size_t, int, int, nccl_comm, cudaStream_t);
typedef int (*fn_bcast)(const void*, void*, size_t, int, int, nccl_comm, cudaStream_t);
typedef int (*fn_destroy)(nccl_comm);
typedef const char* (*fn_errstr)(int);
typedef int (*fn_version)(int*);
constexpr int kNcclInt8 = 0, kNcclFloat32 = 7, kNcclSum = 0, kNcclAvg = 4;

struct NcclApi {
  void* handle = nullptr;
  fn_get_uid get_uid = nullptr; fn_init_rank init_rank = nullptr; fn_allreduce allreduce = nullptr; fn_bcast bcast = nullptr;
  fn_destroy destroy = nullptr; fn_errstr errstr = nullptr; fn_version version = nullptr;
  bool tried = false;
} g_nccl;

bool nccl_load() {
  if (g_nccl.tried) return g_nccl.handle != nullptr;
  g_nccl.tried = true;
  const char* names[3] = {getenv("RD_B200_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
  void* h = nullptr;
  for (int i = 0; i < 3 && !h; ++i) {
    if (!names[i]) continue;
    h = dlopen(names[i], RTLD_NOW | RTLD_NOLOAD | RTLD_GLOBAL);        // the copy a hosting process (PyTorch) already loaded
    if (!h) h = dlopen(names[i], RTLD_NOW | RTLD_GLOBAL);
  }
  if (!h) return false;
  g_nccl.get_uid = (fn_get_uid)dlsym(h, "ncclGetUniqueId");
  g_nccl.init_rank = (fn_init_rank)dlsym(h, "ncclCommInitRank");
  g_nccl.allreduce = (fn_allreduce)dlsym(h, "ncclAllReduce");
  g_nccl.bcast = (fn_bcast)dlsym(h, "ncclBroadcast");
  g_nccl.destroy = (fn_destroy)dlsym(h, "ncclCommDestroy");
  g_nccl.errstr = (fn_errstr)dlsym(h, "ncclGetErrorString");
  g_nccl.version = (fn_version)dlsym(h, "ncclGetVersion");
  if (!g_nccl.get_uid || !g_nccl.init_rank || !g_nccl.allreduce || !g_nccl.bcast || !g_nccl.destroy) return false;
  g_nccl.handle = h;
  return true;
}
const char* nccl_err(int rc) { return g_nccl.errstr ? g_nccl.errstr(rc) : "nccl error"; }
}  // namespace

#define RD_NCCL(ctx, call)                                                                         \
  do {                                                                                             \
    int rc__ = (call);                                                                             \
    if (rc__ != 0) RD_FAIL(ctx, RD_ERR_CUDA, "%s failed: %s", #call, nccl_err(rc__));              \
  } while (0)

extern "C" int rd_ddp_available(rd_ctx* ctx, int* version) {
  (void)ctx;
  if (!nccl_load()) return 0;
  if (version && g_nccl.version) g_nccl.version(version);
  return 1;
}
extern "C" int rd_ddp_unique_id(rd_ctx* ctx, void* id128) {
  if (!nccl_load()) RD_FAIL(ctx, RD_ERR_UNSUPPORTED, "ddp: libnccl.so.2 not found (set RD_B200_NCCL_LIB)");
  nccl_uid uid;
  RD_NCCL(ctx, g_nccl.get_uid(&uid));
  memcpy(id128, &uid, sizeof(uid));
  return RD_OK;
}
extern "C" int rd_ddp_init(rd_ctx* ctx, int world, int rank, const void* id128) {
  if (!nccl_load()) RD_FAIL(ctx, RD_ERR_UNSUPPORTED, "ddp: libnccl.so.2 not found (set RD_B200_NCCL_LIB)");
  if (ctx->nccl_comm) RD_FAIL(ctx, RD_ERR_ARG, "ddp_init: communicator already initialised");
  if (world < 1 || rank < 0 || rank >= world) RD_FAIL(ctx, RD_ERR_ARG, "ddp_init: rank %d of %d", rank, world);
  nccl_uid uid;
  memcpy(&uid, id128, sizeof(uid));
  RD_CUDA(ctx, cudaSetDevice(ctx->device));
  nccl_comm comm = nullptr;
  RD_NCCL(ctx, g_nccl.init_rank(&comm, world, uid, rank));
  ctx->nccl_comm = comm; ctx->ddp_world = world; ctx->ddp_rank = rank;
  return RD_OK;
}
extern "C" int rd_ddp_bucket_allreduce(rd_ctx* ctx, float* grad, int64_t n, int average, rd_stream st) {
  if (!ctx->nccl_comm) RD_FAIL(ctx, RD_ERR_ARG, "ddp_bucket_allreduce: rd_ddp_init first");
  RD_NCCL(ctx, g_nccl.allreduce(grad, grad, (size_t)n, kNcclFloat32, average ? kNcclAvg : kNcclSum, ctx->nccl_comm, (cudaStream_t)st));
  return RD_OK;
}
extern "C" int rd_ddp_broadcast(rd_ctx* ctx, void* buf, int64_t bytes, int root, rd_stream st) {
  if (!ctx->nccl_comm) RD_FAIL(ctx, RD_ERR_ARG, "ddp_broadcast: rd_ddp_init first");
  RD_NCCL(ctx, g_nccl.bcast(buf, buf, (size_t)bytes, kNcclInt8, root, ctx->nccl_comm, (cudaStream_t)st));
  return RD_OK;
}
extern "C" int rd_ddp_finalize(rd_ctx* ctx) {
  if (ctx->nccl_comm) {
    RD_NCCL(ctx, g_nccl.destroy(ctx->nccl_comm));
    ctx->nccl_comm = nullptr;
  }
  return RD_OK;
}
