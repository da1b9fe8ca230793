// gather.cu -- host side of the multi-GPU output gather (see engine.hpp, struct
// Gather, and inverse.cuh, GatherArgs): buffer ownership, CUDA IPC / peer mapping,
// per-block kernel arguments, the root's wait / release launches.
#include "engine.hpp"

#include <algorithm>
#include <cstdlib>

namespace ssr_b200
{

using namespace ssrk;

// one-CTA helper kernels, launched into the programmatic-dependent-launch chain of the
// block loop (their launch latency overlaps the predecessor's tail)
template<typename... KArgs, typename... Args>
static void launch_small(void (*kernel)(KArgs...), int threads, cudaStream_t stream, bool pdl, Args&&... args)
{
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(1); cfg.blockDim = dim3(threads); cfg.dynamicSmemBytes = 0; cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl ? 1 : 0;
  SSR_CUDA(cudaLaunchKernelEx(&cfg, kernel, KArgs(std::forward<Args>(args))...));
}

// ------------------------------------------------------------------ Gather
Gather::Gather(Engine* e_, int world_, int rank_, int root_, const int* outputs_per_rank)
  : e(e_), world(world_), rank(rank_), root(root_)
{
  if (world < 1 || world > kMaxWorld || rank < 0 || rank >= world || root < 0 || root >= world || !outputs_per_rank)
    throw Error(SSR_ERR_INVALID, "ssr_b200: gather: bad world / rank / root");
  for (int g = 0; g < world; ++g)
  {
    if (outputs_per_rank[g] < 0) throw Error(SSR_ERR_INVALID, "ssr_b200: gather: negative output count");
    firsts.push_back(M_total);
    counts.push_back(outputs_per_rank[g]);
    M_total += outputs_per_rank[g];
  }
  if (M_total == 0) throw Error(SSR_ERR_INVALID, "ssr_b200: gather: no outputs");
  e->use_device();
  // SSR_B200_GATHER_TIMEOUT_MS: how long a rank waits for another one before it gives up
  if (const char* v = std::getenv("SSR_B200_GATHER_TIMEOUT_MS"))
    timeout_ns = (unsigned long long)(std::max(1.0, std::atof(v)) * 1e6);
  h_error.resize(16);
  h_error.ptr[0] = 0;
  {
    void* mapped = nullptr;
    SSR_CUDA(cudaHostGetDevicePointer(&mapped, h_error.ptr, 0));
    d_error = static_cast<int*>(mapped);
  }
  if (is_root())
  {
    void* p = nullptr;
    cudaError_t err = cudaMalloc(&p, total_bytes());
    if (err != cudaSuccess)
      throw Error(err == cudaErrorMemoryAllocation ? SSR_ERR_NOMEM : SSR_ERR_CUDA,
          std::string("cudaMalloc (gather buffer): ") + cudaGetErrorString(err));
    base = static_cast<unsigned char*>(p);
    owns = true;
    SSR_CUDA(cudaMemset(base, 0, total_bytes()));
    std::vector<int> ctas(counts);
    ctas[root] = 0;  // the root's own blocks are ordered by its stream
    d_ctas.resize(world);
    SSR_CUDA(cudaMemcpy(d_ctas.ptr, ctas.data(), world * sizeof(int), cudaMemcpyHostToDevice));
  }
}

Gather::~Gather()
{
  cudaSetDevice(e->device);
  cudaStreamSynchronize(e->stream);
  if (owns && base) cudaFree(base);
  else if (ipc_mapped && base) cudaIpcCloseMemHandle(base);
}

void Gather::export_handle(void* handle64) const
{
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
  if (!is_root() || !base) throw Error(SSR_ERR_STATE, "ssr_b200: gather: only the root exports its buffer");
  e->use_device();
  cudaIpcMemHandle_t h;
  SSR_CUDA(cudaIpcGetMemHandle(&h, base));
  std::memcpy(handle64, &h, sizeof(h));
}

void Gather::import_handle(const void* handle64)
{
  if (is_root() || base) throw Error(SSR_ERR_STATE, "ssr_b200: gather: buffer already present");
  e->use_device();
  cudaIpcMemHandle_t h;
  std::memcpy(&h, handle64, sizeof(h));
  void* p = nullptr;
  SSR_CUDA(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
  base = static_cast<unsigned char*>(p);
  ipc_mapped = true;
}

void Gather::import_pointer(void* root_base, int root_device)
{
  if (is_root() || base) throw Error(SSR_ERR_STATE, "ssr_b200: gather: buffer already present");
  if (!root_base) throw Error(SSR_ERR_INVALID, "ssr_b200: gather: null buffer");
  e->use_device();
  if (root_device != e->device)
  {
    int can = 0;
    SSR_CUDA(cudaDeviceCanAccessPeer(&can, e->device, root_device));
    if (!can) throw Error(SSR_ERR_UNSUPPORTED, "ssr_b200: gather: no peer access to the root device");
    cudaError_t err = cudaDeviceEnablePeerAccess(root_device, 0);
    if (err == cudaErrorPeerAccessAlreadyEnabled) cudaGetLastError();
    else SSR_CUDA(err);
  }
  else shares_root_device = true;
  base = static_cast<unsigned char*>(root_base);
}

GatherArgs Gather::next_block()
{
  if (!base) throw Error(SSR_ERR_STATE, "ssr_b200: gather: buffer not connected (import the root's handle first)");
  GatherArgs ga{};
  const uint64_t t = epoch++;
  // InvJob::out holds the offset of the output inside a slot (see Renderer)
  ga.out_offset = (long long)((t & 1) * slot_floats());
  ga.error = d_error;
  ga.timeout_ns = timeout_ns;
  if (!is_root())
  {
    ga.consumed = consumed_ptr();
    ga.need_consumed = t >= 2 ? t - 1 : 0;   // block t-2 (same slot) released
    ga.arrived = arrived_ptr(rank);
  }
  return ga;
}

const float* Gather::wait()
{
  if (!is_root()) throw Error(SSR_ERR_STATE, "ssr_b200: gather: only the root waits");
  if (epoch == 0 || waited == epoch) throw Error(SSR_ERR_STATE, "ssr_b200: gather: no rendered block to wait for");
  e->use_device();
  const uint64_t t = epoch - 1;
  if (world > 1)
  {
    GatherWaitArgs a{arrived_ptr(0), d_ctas.ptr, world, t, d_error, timeout_ns};
    launch_small(gather_wait_kernel, 32, e->stream, e->pdl, a);
    e->count_launch();
  }
  waited = epoch;
  return slot_ptr(int(t & 1));
}

void Gather::release(cudaStream_t s)
{
  if (!is_root()) throw Error(SSR_ERR_STATE, "ssr_b200: gather: only the root releases");
  if (released == waited) return;
  e->use_device();
  released = waited;
  if (world > 1)
  {
    launch_small(gather_release_kernel, 1, s ? s : e->stream, e->pdl, consumed_ptr(), (unsigned long long)released);
    e->count_launch();
  }
}

int Gather::status()
{
  e->use_device();
  e->sync();
  return timed_out() ? 1 : 0;
}

}  // namespace ssr_b200
