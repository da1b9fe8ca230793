// engine.cu -- engine core: device resources, launches, filter sets, Level-1 objects.
#include "engine.hpp"
#include "fft1024.cuh"

#include <algorithm>
#include <cmath>
#include <cstdlib>

namespace ssr_b200
{

using namespace ssrk;

// ------------------------------------------------------------------ MacPlan
void MacPlan::make_work(int slots, int fixed_cost, int waves, const std::vector<double>* weights)
{
  work.host.clear();
  cta_first.host.clear();
  cta_head.host.clear();
  npartials = 0; ncta = 0;
  long total = 0;
  for (auto& g : groups) { total += g.term_end - g.term_begin; g.partial_first = 0; g.nsplit = 0; }
  if (groups.empty()) return;
  // Number of CTAs: one full wave of the resident CTAs.  Every extra wave costs each SM
  // another pipeline fill and another set of partials (waves x (share + fixed cost) only
  // grows), and since shares may straddle group boundaries nothing forces more; small
  // problems use fewer CTAs so that a share is worth its fixed cost.
  const long min_share = std::max<long>(1, fixed_cost / 2);
  long C = std::min<long>(long(slots) * std::max(1, waves), (total + min_share - 1) / min_share);
  C = std::max<long>(C, 1);
  // groups without terms still own one (empty) segment: their accumulators are zeros
  std::vector<int4> segs;
  std::vector<int> first;
  size_t g = 0;
  // share boundaries: equal, or proportional to the calibrated speed of each CTA
  std::vector<long> bound(size_t(C) + 1);
  if (weights && long(weights->size()) == C)
  {
    double sum = 0.0, run = 0.0;
    for (double w : *weights) sum += std::max(w, 1e-9);
    bound[0] = 0;
    for (long c = 0; c < C; ++c)
    {
      run += std::max((*weights)[size_t(c)], 1e-9);
      bound[size_t(c) + 1] = std::max(bound[size_t(c)], std::min<long>(total, long(double(total) * run / sum + 0.5)));
    }
    bound[size_t(C)] = total;
  }
  else
    for (long c = 0; c <= C; ++c) bound[size_t(c)] = total * c / C;
  for (long c = 0; c < C; ++c)
  {
    const long b = bound[size_t(c)], e = bound[size_t(c) + 1];   // share of CTA c in the term table
    first.push_back(int(segs.size()));
    long pos = b;
    while (pos < e)
    {
      while (groups[g].term_end <= pos) ++g;                  // (terms of all groups are back to back)
      const long stop = std::min<long>(e, groups[g].term_end);
      segs.push_back(make_int4(int(pos), int(stop), int(g), 0));
      pos = stop;
    }
  }
  first.push_back(int(segs.size()));
  // partial indices in term order (so a group's partials are consecutive) ...
  for (size_t i = 0; i < segs.size(); ++i)
  {
    Group& grp = groups[size_t(segs[i].z)];
    if (grp.nsplit == 0) grp.partial_first = int(i);
    grp.nsplit++;
    segs[i].z = int(i);
  }
  npartials = int(segs.size());
  // ... and the longest segment of every CTA first: the bulk-copy kernel overlaps its
  // first segment's p != 0 terms with the block's forward transform
  for (long c = 0; c < C; ++c)
  {
    const int lo = first[size_t(c)], hi = first[size_t(c) + 1];
    int best = lo;
    for (int i = lo + 1; i < hi; ++i)
      if (segs[size_t(i)].y - segs[size_t(i)].x > segs[size_t(best)].y - segs[size_t(best)].x) best = i;
    if (best != lo) std::swap(segs[size_t(lo)], segs[size_t(best)]);
  }
  work.host = segs;
  cta_first.host = first;
  for (long c = 0; c < C; ++c)
  {
    const int lo = first[size_t(c)], hi = first[size_t(c) + 1];
    cta_head.host.push_back(make_int4(lo, hi, 0, 0));
    cta_head.host.push_back(lo < hi ? segs[size_t(lo)] : make_int4(0, 0, 0, 0));
  }
  ncta = int(C);
}

// ------------------------------------------------------------------ launches
// pdl: programmatic dependent launch -- the kernel may be scheduled while the
// previous kernel of the stream drains (it starts with grid_dependency_wait()),
// which hides most of the launch latency between the 3-4 small kernels of a block.
template<typename... KArgs, typename... Args>
static void launch_kernel(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                          bool pdl, Args&&... args)
{
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl ? 1 : 0;
  SSR_CUDA(cudaLaunchKernelEx(&cfg, kernel, KArgs(std::forward<Args>(args))...));
}

// the same with thread-block clusters of `cluster` CTAs along x (grid.x is a multiple of it)
template<typename... KArgs, typename... Args>
static void launch_cluster_kernel(void (*kernel)(KArgs...), dim3 grid, dim3 block, int cluster, size_t smem,
                                  cudaStream_t stream, bool pdl, Args&&... args)
{
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = unsigned(cluster); attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl ? 2 : 1;
  SSR_CUDA(cudaLaunchKernelEx(&cfg, kernel, KArgs(std::forward<Args>(args))...));
}

// ------------------------------------------------------------------ Engine
// float4 per thread and 1024-bin tiles (gridDim.y) of the register-staged MAC for block
// size B (B / 2 float4 per partition spectrum; any multiple of 8)
static int mac_nv(int B) { return B / 2 > kThreads ? 2 : 1; }
static int mac_ktiles(int B) { return std::max(1, (B / 2 + 2 * kThreads - 1) / (2 * kThreads)); }
static bool is_pow2(int B) { return (B & (B - 1)) == 0; }

template<int MT, int NV, int ST = 2, int MINB = 1>
static void* mac_fn() { return (void*)mac_kernel<MT, NV, ST, MINB>; }

// variant (ssr_engine_config::mac_variant), experiments on the Generic kernel
// (profiles/r1_mac_variants.md):
//   0 default = bulk-copy (TMA) staged mac_tma_kernel with 5 stages at B >= 1024
//     (6.78 TB/s on cfg5 under the power cap where the register-staged kernel
//     reaches 6.37), the register-staged kernel below otherwise
//   6 / 7 = mac_tma_kernel with 4 / 5 stages
//   2 = register-staged: 2-stage pipeline, 2 CTAs / SM (128 registers) -- the
//     default before the bulk-copy kernel existed
//   1 = 3 stages, 4 = 4 stages (1 CTA / SM), 5 = 2 stages at 1 CTA / SM,
//   3 = variant 2 without L2 eviction hints
static void* mac_kernel_ptr(int MT, int NV, int variant)
{
  if (NV == 2 && MT == 4)
  {
    if (variant == 1) return mac_fn<4, 2, 3, 1>();
    if (variant == 4) return mac_fn<4, 2, 4, 1>();
    if (variant == 5) return mac_fn<4, 2, 2, 1>();
    return mac_fn<4, 2, 2, 2>();
  }
  if (NV == 2)
  {
    if (MT == 1) return mac_fn<1, 2>();
    return mac_fn<2, 2>();
  }
  if (MT == 1) return mac_fn<1, 1>();
  if (MT == 2) return mac_fn<2, 1>();
  return mac_fn<4, 1>();
}

Engine::Engine(const ssr_engine_config& cfg)
  : device(cfg.device), B(cfg.block_size), sample_rate(cfg.sample_rate)
  , mac_variant(cfg.mac_variant), stream(nullptr), own_stream(false)
{
  if (mac_variant == 0)
    if (const char* v = std::getenv("SSR_B200_MAC_VARIANT")) mac_variant = std::atoi(v);  // experiments
  // B = 1024: one-warp-per-transform register FFT (fft1024.cuh); SSR_B200_FFT=generic
  // selects the shared-memory Stockham kernels instead (A/B experiments)
  const char* fft_env = std::getenv("SSR_B200_FFT");
  fft1024 = (B == 1024) && !(fft_env && std::string(fft_env) == "generic");
  fft1024_always = fft_env && std::string(fft_env) == "fft1024";  // also for small launches (tests, A/B)
  if (B <= 0 || B % 8 != 0)
    throw Error(SSR_ERR_BLOCK_SIZE, "Convolver: block size must be a multiple of 8!");
  if (B > kMaxBlock)
    throw Error(SSR_ERR_UNSUPPORTED, "ssr_b200: block size must be a multiple of 8 in [8, 8192]");
  int ndev = 0;
  cudaError_t err = cudaGetDeviceCount(&ndev);
  if (err != cudaSuccess || ndev == 0)
    throw Error(SSR_ERR_CUDA, std::string("ssr_b200: no CUDA device usable (")
        + cudaGetErrorString(err) + "); this engine has no CPU fallback");
  if (device < 0 || device >= ndev) throw Error(SSR_ERR_INVALID, "ssr_b200: bad device ordinal");
  use_device();
  cudaDeviceProp prop;
  SSR_CUDA(cudaGetDeviceProperties(&prop, device));
  sm_count = prop.multiProcessorCount;
  if (cfg.stream) stream = static_cast<cudaStream_t>(cfg.stream);
  else { SSR_CUDA(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking)); own_stream = true; }

  // twiddles exp(-2 pi i k / 2B), computed in double
  const int n = 2 * B;
  std::vector<float2> htw(n);
  for (int k = 0; k < n; ++k)
  {
    const double a = -2.0 * M_PI * double(k) / double(n);
    htw[k] = make_float2(float(std::cos(a)), float(std::sin(a)));
  }
  tw.resize(n);
  SSR_CUDA(cudaMemcpy(tw.ptr, htw.data(), n * sizeof(float2), cudaMemcpyHostToDevice));

  // raised-cosine crossfade, same float expression as the reference
  // (apf/apf/math.h:254-259 with period 2B, apf/apf/combine_channels.h:481-497)
  std::vector<float> w(B + 1), fo(B), fi(B);
  const float period = float(2 * B);
  const float pi_f = float(M_PI);
  for (int i = 0; i <= B; ++i) w[i] = std::cos(float(i) * 2 * pi_f / period) * 0.5f + 0.5f;
  for (int i = 0; i < B; ++i) { fo[i] = w[i]; fi[i] = w[B - i]; }
  fade_out.resize(B); fade_in.resize(B);
  SSR_CUDA(cudaMemcpy(fade_out.ptr, fo.data(), B * sizeof(float), cudaMemcpyHostToDevice));
  SSR_CUDA(cudaMemcpy(fade_in.ptr, fi.data(), B * sizeof(float), cudaMemcpyHostToDevice));

  zero_part.resize(B);
  SSR_CUDA(cudaMemset(zero_part.ptr, 0, B * sizeof(float2)));
  d_dyn_stats.resize(2);
  SSR_CUDA(cudaMemset(d_dyn_stats.ptr, 0, 2 * sizeof(unsigned long long)));

  d_xring.resize(kMaxInputs); d_xparts.resize(kMaxInputs); d_heads.resize(kMaxInputs);
  SSR_CUDA(cudaMemset(d_xring.ptr, 0, kMaxInputs * sizeof(void*)));
  SSR_CUDA(cudaMemset(d_xparts.ptr, 0, kMaxInputs * sizeof(int)));
  SSR_CUDA(cudaMemset(d_heads.ptr, 0, kMaxInputs * sizeof(int)));

  const size_t fft_smem = size_t(2) * B * sizeof(float2);
  if (mac_tma())
  {
    SSR_CUDA(cudaFuncSetAttribute(mac_tma_batch_kernel<4, Input::kTimeBatch, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize,
          int(mac_tma_batch_smem_bytes<4, Input::kTimeBatch, 3>())));
    SSR_CUDA(cudaFuncSetAttribute(mac_tma_kernel<4, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize,
          int(mac_tma_smem_bytes<4, 4>())));
    SSR_CUDA(cudaFuncSetAttribute(mac_tma_kernel<4, 5>, cudaFuncAttributeMaxDynamicSharedMemorySize,
          int(mac_tma_smem_bytes<4, 5>())));
  }
  if (const char* v = std::getenv("SSR_B200_DYN_TMA")) dyn_tma_enabled = std::atoi(v) != 0;   // A/B experiments
  if (dyn_tma())
  {
    SSR_CUDA(cudaFuncSetAttribute(mac_dyn_tma_kernel<5>, cudaFuncAttributeMaxDynamicSharedMemorySize,
          int(mac_dyn_tma_smem_bytes<5>())));
    SSR_CUDA(cudaFuncSetAttribute(inv_dyn_cluster_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(fft_smem * 4)));
    SSR_CUDA(cudaFuncSetAttribute(fwd_fft_overlap_kernel<2>, cudaFuncAttributePreferredSharedMemoryCarveout,
          int(cudaSharedmemCarveoutMaxShared)));
    if (2 * fft_smem > 48 * 1024)
      SSR_CUDA(cudaFuncSetAttribute(fwd_fft_overlap_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(fft_smem)));
    // cluster size of the inverse kernel: 16 CTAs (non-portable) when the device schedules them, else 8
    dyn_cluster = 8;
    if (cudaFuncSetAttribute(inv_dyn_cluster_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) == cudaSuccess)
    {
      cudaLaunchConfig_t qc{};
      qc.gridDim = dim3(32); qc.blockDim = dim3(kThreads * 3); qc.dynamicSmemBytes = fft_smem * 4;
      cudaLaunchAttribute qa[1];
      qa[0].id = cudaLaunchAttributeClusterDimension;
      qa[0].val.clusterDim.x = 16; qa[0].val.clusterDim.y = 1; qa[0].val.clusterDim.z = 1;
      qc.attrs = qa; qc.numAttrs = 1;
      int nclusters = 0;
      if (cudaOccupancyMaxActiveClusters(&nclusters, inv_dyn_cluster_kernel, &qc) == cudaSuccess && nclusters >= 2)
        dyn_cluster = 16;
    }
    (void)cudaGetLastError();
    if (const char* v = std::getenv("SSR_B200_DYN_CLUSTER"))
    {
      const int c = std::atoi(v);
      if (c == 1 || c == 2 || c == 4 || c == 8 || c == 16) dyn_cluster = c;
    }
    while (dyn_cluster > 1 && (B / 2) % dyn_cluster != 0) dyn_cluster /= 2;
  }
  // the overlapped forward kernel shares its SM with a MAC CTA (205 KiB of shared memory):
  // both must run under the same (maximal) shared-memory carve-out to be co-resident
  SSR_CUDA(cudaFuncSetAttribute(fwd_fft_overlap_kernel<0>, cudaFuncAttributePreferredSharedMemoryCarveout,
        int(cudaSharedmemCarveoutMaxShared)));
  // (the generated-terms MAC of Binaural / BRS keeps the default carve-out -- with the
  // maximal one it loses 20 % (profiles/r2_ab) -- and so does the forward kernel next to it)
  if (const char* v = std::getenv("SSR_B200_PDL")) pdl = std::atoi(v) != 0;  // A/B experiments
  if (const char* v = std::getenv("SSR_B200_FWD_OVERLAP")) fwd_overlap = std::atoi(v) != 0;
  if (const char* v = std::getenv("SSR_B200_MAC_WAVES")) mac_waves = std::max(0, std::atoi(v));
  if (const char* v = std::getenv("SSR_B200_PREREDUCE_MIN")) prereduce_min = std::max(1, std::atoi(v));
  if (const char* v = std::getenv("SSR_B200_MAC_CALIBRATE")) mac_calibrate = std::atoi(v) != 0;
  if (const char* v = std::getenv("SSR_B200_INV_TRIGGER")) inv_trigger = std::atoi(v) != 0;   // A/B experiments
  if (const char* v = std::getenv("SSR_B200_EARLY_TRIGGER")) early_trigger = std::atoi(v) != 0;   // A/B experiments (profiles/r2_ab)
  if (const char* v = std::getenv("SSR_B200_TRACE_ISSUE")) trace_mode = std::atoi(v) != 0 ? 1 : 0;   // scripts/mac_trace.py
  if (mac_tma())
  {
    mac_trace.resize(4096 * 4);
    SSR_CUDA(cudaMemset(mac_trace.ptr, 0, mac_trace.count * sizeof(unsigned long long)));
  }
  fwd_overlap = fwd_overlap && pdl;
  if (const char* v = std::getenv("SSR_B200_BLOCK_OVERLAP")) block_overlap = std::atoi(v) != 0;   // A/B (profiles/r2c_ab)
  // the inverse kernel may only commit the ring heads first if the MAC in front of it is complete when it starts
  block_overlap = block_overlap && fwd_overlap && !early_trigger;
  if (const char* v = std::getenv("SSR_B200_LEVEL1")) level1_fused = std::string(v) != "unfused";  // A/B experiments
  if (!is_pow2(B)) level1_fused = false;   // the fused Level-1 launch is instantiated for power-of-two spectra
  // 4096 < B <= 8192: the two transform buffers take 128 KiB; kernels that also stage the twiddle table or
  // run several transforms side by side have "large" variants without (fwd_fft_large_kernel, inv_fft_large_kernel)
  if (large()) level1_fused = false;
  if (large())
  {
    SSR_CUDA(cudaFuncSetAttribute(fwd_fft_large_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(fft_smem)));
    SSR_CUDA(cudaFuncSetAttribute(inv_fft_large_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(fft_smem)));
    SSR_CUDA(cudaFuncSetAttribute(fwd_fft_batch_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(fft_smem)));
    SSR_CUDA(cudaFuncSetAttribute(fwd_fft_overlap_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(fft_smem)));
    SSR_CUDA(cudaFuncSetAttribute(fwd_fft_overlap_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(fft_smem)));
  }
  else if (2 * fft_smem > 48 * 1024)
  {
    SSR_CUDA(cudaFuncSetAttribute(level1_fused_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(2 * fft_smem)));
    SSR_CUDA(cudaFuncSetAttribute(level1_fused_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(2 * fft_smem)));
    SSR_CUDA(cudaFuncSetAttribute(fwd_fft_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(2 * fft_smem)));
    SSR_CUDA(cudaFuncSetAttribute(fwd_fft_batch_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(fft_smem)));
    SSR_CUDA(cudaFuncSetAttribute(fwd_fft_overlap_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(fft_smem)));
    SSR_CUDA(cudaFuncSetAttribute(fwd_fft_overlap_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(fft_smem)));
  }
  // the inverse kernel runs up to kInvMaxGroups transforms side by side (3 x 16 KiB
  // at B = 1024, which with its static shared memory already exceeds 48 KiB)
  // (+ one staged twiddle table per CTA)
  if (!large())
    SSR_CUDA(cudaFuncSetAttribute(inv_fft_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(fft_smem * 2)));
  if (inv_groups(2) == 2)
    SSR_CUDA(cudaFuncSetAttribute(inv_fft_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(fft_smem * 3)));
  if (inv_groups(3) == 3)
    SSR_CUDA(cudaFuncSetAttribute(inv_fft_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(fft_smem * 4)));
  SSR_CUDA(cudaDeviceSynchronize());
}

Engine::~Engine()
{
  cudaSetDevice(device);
  cudaStreamSynchronize(stream);
  for (auto& p : pending) for (auto& ev : p.ev) cudaEventDestroy(ev);
  for (auto& p : free_events) for (auto& ev : p.ev) cudaEventDestroy(ev);
  if (own_stream) cudaStreamDestroy(stream);
}

int Engine::register_input(Input* in)
{
  int id = -1;
  for (size_t i = 0; i < inputs.size(); ++i) if (!inputs[i]) { id = int(i); break; }
  if (id < 0)
  {
    if (int(inputs.size()) >= kMaxInputs) throw Error(SSR_ERR_NOMEM, "ssr_b200: too many inputs");
    id = int(inputs.size());
    inputs.push_back(nullptr);
  }
  inputs[id] = in;
  const float4* ring = reinterpret_cast<const float4*>(in->ring.ptr);
  const int parts = in->slots, zero = 0;
  SSR_CUDA(cudaMemcpyAsync(d_xring.ptr + id, &ring, sizeof(ring), cudaMemcpyHostToDevice, stream));
  SSR_CUDA(cudaMemcpyAsync(d_xparts.ptr + id, &parts, sizeof(int), cudaMemcpyHostToDevice, stream));
  SSR_CUDA(cudaMemcpyAsync(d_heads.ptr + id, &zero, sizeof(int), cudaMemcpyHostToDevice, stream));
  sync();  // the sources above are stack variables
  return id;
}

void Engine::unregister_input(int id)
{
  if (id >= 0 && id < int(inputs.size())) inputs[id] = nullptr;
}

// bulk-copy (TMA) staged MAC for the four-row (Generic) plans at B >= 1024
bool Engine::mac_tma() const { return (mac_variant == 0 || mac_variant == 6 || mac_variant == 7) && B >= 1024 && is_pow2(B); }

// Fixed cost of a CTA in units of one term, for MacPlan::make_work: staging and the
// partial spectra it writes (and the inverse kernel sums); the bulk-copy kernel also
// fills its 5-stage ring before the first FMA.  SSR_B200_SPLIT_COST overrides (experiments).
int Engine::split_cost(int MT) const
{
  if (const char* v = std::getenv("SSR_B200_SPLIT_COST")) return std::max(0, std::atoi(v));
  return (mac_tma() && MT == 4) ? kSplitCostTma : 2;
}

int Engine::mac_slots(int MT) const
{
  // one CTA per SM (shared-memory bound); B > 1024: every share runs as mac_ktiles(B) CTAs, one per
  // 1024-bin tile of the spectrum, so fewer shares fill the same single wave
  if (mac_tma() && MT == 4) return std::max(1, sm_count / mac_ktiles(B));
  int occ = 0;
  cudaError_t err = cudaOccupancyMaxActiveBlocksPerMultiprocessor(
      &occ, mac_kernel_ptr(MT, mac_nv(B), mac_variant), kThreads, 0);
  if (err != cudaSuccess || occ < 1) occ = 1;
  return sm_count * occ;
}

void Engine::launch_fwd(const FwdJob* d_jobs, int njobs, const FwdRingUpdate* ring_update, bool overlap,
                        const float* in_base, bool late_wait, bool dyn_variant)
{
  if (njobs <= 0) return;
  FwdRingUpdate ru{nullptr, nullptr, nullptr, 0};
  // One warp per transform wins on throughput (2.2x at filter load), one CTA per
  // transform on latency; the block loop rarely has enough transforms to fill
  // the machine (profiles/r1_fft_kernels.md)
  if (overlap)
  {
    if (ring_update) { ru = *ring_update; ring_update = nullptr; }
    if ((ru.ctl || dyn_variant) && dyn_tma())
      launch_kernel(fwd_fft_overlap_kernel<2>, dim3(njobs), dim3(kThreads), size_t(2) * B * sizeof(float2), stream, pdl,
          d_jobs, B, static_cast<const float2*>(tw.ptr), ru, in_base, late_wait ? 1 : 0);
    else if (ru.ctl)
      launch_kernel(fwd_fft_overlap_kernel<1>, dim3(njobs), dim3(kThreads), size_t(2) * B * sizeof(float2), stream, pdl,
          d_jobs, B, static_cast<const float2*>(tw.ptr), ru, in_base, 0);
    else
      launch_kernel(fwd_fft_overlap_kernel<0>, dim3(njobs), dim3(kThreads), size_t(2) * B * sizeof(float2), stream, pdl,
          d_jobs, B, static_cast<const float2*>(tw.ptr), ru, in_base, late_wait ? 1 : 0);
  }
  else if (fft1024 && !in_base && (njobs >= kFft1024MinJobs || fft1024_always))
  {
    const int ctas = (njobs + kWarpsPerCta - 1) / kWarpsPerCta;
    fwd_fft1024_kernel<<<ctas, kWarpsPerCta * 32, 0, stream>>>(d_jobs, njobs, tw.ptr);
  }
  else if (large())
  {
    if (ring_update) { ru = *ring_update; ring_update = nullptr; }
    fwd_fft_large_kernel<<<njobs, kThreads, size_t(2) * B * sizeof(float2), stream>>>(d_jobs, B, tw.ptr, ru, in_base);
  }
  else
  {
    const size_t smem = size_t(4) * B * sizeof(float2);  // data + staged twiddles
    if (ring_update) { ru = *ring_update; ring_update = nullptr; }  // folded into this launch
    fwd_fft_kernel<<<njobs, kThreads, smem, stream>>>(d_jobs, B, tw.ptr, ru, in_base);
  }
  SSR_CUDA(cudaGetLastError());
  count_launch();
  if (cur) cur->fft_launches++;
  if (ring_update)  // the warp-per-transform kernel ran: update the filter ring separately
    launch_dyn_ring_update(ring_update->ctl, ring_update->upd, ring_update->ring, 2 * njobs, ring_update->R);
}

void Engine::launch_fwd_batch(const FwdBatch& batch, int nparts, int nchannels)
{
  if (nparts <= 0 || nchannels <= 0) return;
  if (fft1024)
  {
    const long total = long(nparts) * nchannels;
    const long ctas = (total + kWarpsPerCta - 1) / kWarpsPerCta;
    fwd_fft1024_batch_kernel<<<unsigned(ctas), kWarpsPerCta * 32, 0, stream>>>(batch, nparts, total, tw.ptr);
    SSR_CUDA(cudaGetLastError());
    return;
  }
  const size_t smem = size_t(2) * B * sizeof(float2);
  // grid (partition, channel lo, channel hi): gridDim.y is limited to 65535
  const int gy = std::min(nchannels, 32768);
  for (int c0 = 0; c0 < nchannels; c0 += gy * 65535)
  {
    FwdBatch b = batch;
    const int n = std::min(nchannels - c0, gy * 65535);
    b.first_channel += c0;
    b.id_offset += uint64_t(c0);
    if (!b.synth) b.time += long(c0) * b.channel_stride;
    // full rows of gy channels, then the remainder
    const int gz = n / gy;
    if (gz > 0)
    {
      fwd_fft_batch_kernel<<<dim3(nparts, gy, gz), kThreads, smem, stream>>>(b, B, tw.ptr);
      SSR_CUDA(cudaGetLastError());
    }
    const int rem = n - gz * gy;
    if (rem > 0)
    {
      FwdBatch r = b;
      r.first_channel += gz * gy;
      r.id_offset += uint64_t(gz) * gy;
      if (!r.synth) r.time += long(gz) * gy * r.channel_stride;
      fwd_fft_batch_kernel<<<dim3(nparts, rem, 1), kThreads, smem, stream>>>(r, B, tw.ptr);
      SSR_CUDA(cudaGetLastError());
    }
  }
}

void Engine::ensure_partials(size_t nspectra)
{
  const size_t need = nspectra * size_t(B / 2);
  if (need > partials.count)
  {
    sync();  // previous launches may still read the old buffer
    partials.resize(need + need / 4);
  }
}

void Engine::launch_mac_fn(void* fn, dim3 grid, void** args, int threads, size_t smem)
{
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid; cfg.blockDim = dim3(threads); cfg.dynamicSmemBytes = smem; cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl ? 1 : 0;
  SSR_CUDA(cudaLaunchKernelExC(&cfg, fn, args));
}

void Engine::launch_mac(const MacPlan& plan, int wtab_offset, int partial_offset, const float* weights,
                        bool overlap)
{
  const int nwork = plan.work.uploaded ? plan.ncta : 0;
  if (nwork == 0) return;
  MacParams p{};
  p.term_meta = plan.meta.device();
  p.term_h = plan.hptr.device();
  p.work = plan.work.device();
  p.cta_first = plan.cta_first.device();
  p.cta_head = plan.cta_head.device();
  p.head_bias = overlap ? 1 : 0;
  if (mac_trace.ptr) p.trace = mac_trace.ptr;
  p.trace_mode = trace_mode;
  p.early_trigger = early_trigger ? 1 : 0;
  p.wtab = weights ? weights : wtab.device();
  p.xring = d_xring.ptr;
  p.xparts = d_xparts.ptr;
  p.heads = d_heads.ptr;
  p.partials = partials.ptr;
  p.B = B;
  p.wtab_offset = wtab_offset;
  p.partial_offset = partial_offset;
  p.l2_hints = (mac_variant == 3) ? 0 : 1;
  void* args[] = { &p };
  dim3 grid(nwork, mac_ktiles(B), 1);
  stats.last_mac_grid = nwork * mac_ktiles(B);
  if (mac_tma() && plan.MT == 4)
  {
    const bool five = mac_variant != 6;
    void* tfn = five ? (void*)mac_tma_kernel<4, 5> : (void*)mac_tma_kernel<4, 4>;
    const size_t smem = five ? mac_tma_smem_bytes<4, 5>() : mac_tma_smem_bytes<4, 4>();
    launch_mac_fn(tfn, grid, args, kTmaThreads, smem);
  }
  else
  {
    void* fn = mac_kernel_ptr(plan.MT, mac_nv(B), mac_variant);
    launch_mac_fn(fn, grid, args);
  }
  count_launch();
  if (cur) cur->mac_launches++;
}

// generated-terms MAC on the bulk-copy ring + cluster inverse (Binaural / BRS at B = 1024 / 2048: three
// inverse groups side by side must fit the shared memory)
bool Engine::dyn_tma() const { return dyn_tma_enabled && B >= 1024 && is_pow2(B) && inv_groups(3) == 3; }

int Engine::dyn_mac_slots() const
{
  if (dyn_tma()) return sm_count;   // one CTA per SM (shared-memory bound)
  int occ = 0;
  void* fn = mac_nv(B) == 2 ? (void*)mac_kernel<2, 2, 2, 2, true> : (void*)mac_kernel<2, 1, 2, 2, true>;
  cudaError_t err = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, fn, kThreads, 0);
  if (err != cudaSuccess || occ < 1) occ = 1;
  return sm_count * occ;
}

// thread groups (= accumulator entries transformed side by side) of an inverse
// launch: as many as the jobs have entries, within the shared-memory budget
int Engine::inv_groups(int max_entries) const
{
  const size_t per_group = size_t(2) * B * sizeof(float2);  // the twiddle table takes as much again
  if (large()) return 1;
  int g = std::max(1, std::min(max_entries, kInvMaxGroups));
  while (g > 1 && per_group * (g + 1) > size_t(200) * 1024) --g;
  return g;
}

void Engine::launch_inv_kernel(int groups, int njobs, const InvPlan& plan, const float2* red,
                               const GatherArgs& ga, const int* kind_counts, const HeadCommit& hc)
{
  const size_t smem = size_t(2) * B * sizeof(float2) * (groups + 1);
  const float2* part = reinterpret_cast<const float2*>(partials.ptr);
  if (hc.early_trigger == 2) njobs += 1;   // the CTA that commits the ring heads (ssrk::inv_prologue)
  if (large())
    launch_kernel(inv_fft_large_kernel, dim3(njobs), dim3(kThreads), size_t(2) * B * sizeof(float2), stream, pdl,
        plan.jobs.device(), plan.entries.device(), part, red, B, tw.ptr, fade_out.ptr, fade_in.ptr, ga, kind_counts, hc);
  else if (groups == 3)
    launch_kernel(inv_fft_kernel<3>, dim3(njobs), dim3(kThreads * 3), smem, stream, pdl, plan.jobs.device(),
        plan.entries.device(), part, red, B, tw.ptr, fade_out.ptr, fade_in.ptr, ga, kind_counts, hc);
  else if (groups == 2)
    launch_kernel(inv_fft_kernel<2>, dim3(njobs), dim3(kThreads * 2), smem, stream, pdl, plan.jobs.device(),
        plan.entries.device(), part, red, B, tw.ptr, fade_out.ptr, fade_in.ptr, ga, kind_counts, hc);
  else
    launch_kernel(inv_fft_kernel<1>, dim3(njobs), dim3(kThreads), smem, stream, pdl, plan.jobs.device(),
        plan.entries.device(), part, red, B, tw.ptr, fade_out.ptr, fade_in.ptr, ga, kind_counts, hc);
}

void Engine::launch_level1_fused(const FusedParams& p)
{
  const size_t smem = size_t(4) * B * sizeof(float2);
  const int nv = std::max(1, (B / 2) / kThreads);
  if (nv >= 8) level1_fused_kernel<8><<<1, kThreads, smem, stream>>>(p, tw.ptr);
  else if (nv == 4) level1_fused_kernel<4><<<1, kThreads, smem, stream>>>(p, tw.ptr);
  else if (nv == 2) level1_fused_kernel<2><<<1, kThreads, smem, stream>>>(p, tw.ptr);
  else level1_fused_kernel<1><<<1, kThreads, smem, stream>>>(p, tw.ptr);
  SSR_CUDA(cudaGetLastError());
}

void Engine::launch_dyn_ring_update(const DynCtl* ctl, const DynRingEntry* upd, DynRingEntry* ring, int rows, int R)
{
  if (rows <= 0) return;
  dyn_ring_update_kernel<<<(rows + 127) / 128, 128, 0, stream>>>(ctl, upd, ring, rows, R);
  SSR_CUDA(cudaGetLastError());
}

void Engine::launch_mac_batch(const MacPlan& plan, int wtab_offset)
{
  const int nwork = plan.work.uploaded ? plan.ncta : 0;
  if (nwork == 0) return;
  if (!mac_tma() || plan.MT != 4) throw Error(SSR_ERR_UNSUPPORTED, "ssr_b200: time-batched MAC needs the four-row bulk-copy plan (B >= 1024)");
  MacParams p{};
  p.term_meta = plan.meta.device();
  p.term_h = plan.hptr.device();
  p.work = plan.work.device();
  p.cta_first = plan.cta_first.device();
  p.wtab = wtab.device();
  p.xring = d_xring.ptr;
  p.xparts = d_xparts.ptr;
  p.heads = d_heads.ptr;
  p.partials = partials.ptr;
  p.B = B;
  p.wtab_offset = wtab_offset;
  p.partial_offset = 0;
  p.l2_hints = 1;
  void* args[] = { &p };
  dim3 grid(nwork, mac_ktiles(B), 1);
  launch_mac_fn((void*)mac_tma_batch_kernel<4, Input::kTimeBatch, 3>, grid, args, kTmaThreads,
      mac_tma_batch_smem_bytes<4, Input::kTimeBatch, 3>());
  count_launch();
  if (cur) cur->mac_launches++;
}

void Engine::launch_mac_dyn(const DynTables& dyn, int G, const float* weights, bool overlap)
{
  MacParams p{};
  p.head_bias = overlap ? 1 : 0;
  p.wtab = weights;
  p.xring = d_xring.ptr;
  p.xparts = d_xparts.ptr;
  p.heads = d_heads.ptr;
  p.partials = partials.ptr;
  p.B = B;
  p.l2_hints = (mac_variant == 3) ? 0 : 1;
  p.zero = reinterpret_cast<const float4*>(zero_part.ptr);
  p.dyn = dyn;
  p.dyn.stats = cur ? d_dyn_stats.ptr : nullptr;
  p.trace_mode = trace_mode;
  p.early_trigger = early_trigger ? 1 : 0;
  if (dyn_tma() && mac_trace.ptr && size_t(G) * 4 <= mac_trace.count) p.trace = mac_trace.ptr;
  void* args[] = { &p };
  dim3 grid(G, mac_ktiles(B), 1);
  stats.last_mac_grid = G * mac_ktiles(B);
  if (dyn_tma())
    launch_mac_fn((void*)mac_dyn_tma_kernel<5>, grid, args, kTmaThreads, mac_dyn_tma_smem_bytes<5>());
  else
  {
    void* fn = mac_nv(B) == 2 ? (void*)mac_kernel<2, 2, 2, 2, true> : (void*)mac_kernel<2, 1, 2, 2, true>;
    launch_mac_fn(fn, grid, args);
  }
  count_launch();
  if (cur) cur->mac_launches++;
}

// reduce the generated-terms partials per (ear, kind), then one inverse CTA per ear
void Engine::launch_ctl_upload(const void* src_host, void* dst, int n16, unsigned long long* counter,
                               unsigned long long* ack, const FwdRingUpdate* ring_update, int ring_rows)
{
  const FwdRingUpdate ru = ring_update ? *ring_update : FwdRingUpdate{nullptr, nullptr, nullptr, 0};
  launch_kernel(ctl_upload_kernel, dim3(1), dim3(kThreads), 0, stream, pdl, static_cast<const uint4*>(src_host),
      static_cast<uint4*>(dst), n16, counter, static_cast<volatile unsigned long long*>(ack), ru, ring_rows,
      ring_update ? 1 : 0);
  count_launch();
}

void Engine::launch_inv_dyn(const InvPlan& plan, const DynCtl* ctl, int Pmax, int G, const HeadCommit& hc,
                            bool early_trigger)
{
  const int njobs = int(plan.jobs.uploaded);
  if (njobs == 0) return;
  if (dyn_tma())
  {
    // one cluster per ear sums the [G][kind][ear] partial spectra and transforms (no reduction launch)
    // (block overlap: one extra cluster commits the ring heads first, see the kernel)
    launch_cluster_kernel(inv_dyn_cluster_kernel, dim3((njobs + (hc.early_trigger == 2 ? 1 : 0)) * dyn_cluster), dim3(kThreads * 3), dyn_cluster,
        size_t(2) * B * sizeof(float2) * 4, stream, pdl, plan.jobs.device(), plan.entries.device(),
        reinterpret_cast<const float2*>(partials.ptr), G, B, static_cast<const float2*>(tw.ptr),
        static_cast<const float*>(fade_out.ptr), static_cast<const float*>(fade_in.ptr), ctl->n, hc,
        (trace_mode == 1 && mac_trace.ptr) ? mac_trace.ptr + 8192 : static_cast<unsigned long long*>(nullptr),
        hc.early_trigger == 2 ? 2 : (early_trigger ? 1 : 0));
    count_launch();
    if (cur) cur->ifft_launches++;
    return;
  }
  const size_t need = size_t(6) * (B / 2);
  if (need > reduced.count) { sync(); reduced.resize(need); }
  dim3 grid(6, (B / 2 + 31) / 32, 1);
  launch_kernel(dyn_reduce_kernel, grid, dim3(kThreads), 0, stream, pdl, ctl, Pmax, G,
      static_cast<const float4*>(partials.ptr), reduced.ptr, B / 2);
  SSR_CUDA(cudaGetLastError());
  count_launch();
  if (cur) cur->ifft_launches++;
  launch_inv_kernel(inv_groups(3), njobs, plan, reinterpret_cast<const float2*>(reduced.ptr), GatherArgs{}, ctl->n, hc);
  SSR_CUDA(cudaGetLastError());
  count_launch();
  if (cur) cur->ifft_launches++;
}

void Engine::launch_inv(const InvPlan& plan, const GatherArgs& ga, const HeadCommit& hc)
{
  const int njobs = int(plan.jobs.uploaded);
  if (njobs == 0)
  {
    if (hc.n) throw Error(SSR_ERR_STATE, "ssr_b200: an overlapped forward launch needs an inverse launch to commit the ring heads");
    return;
  }
  // few output blocks with deep splits: reduce the partials in parallel first
  const int nentries = int(plan.entries.uploaded);
  int max_count = 0;
  for (auto& en : plan.entries.host) max_count = std::max(max_count, en.part_count);
  const bool pre_reduce = nentries > 0 && max_count >= prereduce_min && njobs < 2 * sm_count;
  const float2* red = nullptr;
  if (pre_reduce)
  {
    const size_t need = size_t(nentries) * (B / 2);
    if (need > reduced.count) { sync(); reduced.resize(need + need / 2); }
    dim3 grid(nentries, (B / 2 + 31) / 32, 1);
    launch_kernel(reduce_partials_kernel, grid, dim3(kThreads), 0, stream, pdl,
        static_cast<const InvEntry*>(plan.entries.device()), static_cast<const float4*>(partials.ptr),
        reduced.ptr, B / 2);
    SSR_CUDA(cudaGetLastError());
    count_launch();
  if (cur) cur->ifft_launches++;
    red = reinterpret_cast<const float2*>(reduced.ptr);
  }
  const bool gathering = ga.out_offset != 0 || ga.consumed || ga.arrived || ga.host_flag;
  if (fft1024 && !gathering && !hc.n && (njobs >= kFft1024MinJobs || fft1024_always))
  {
    const int ctas = (njobs + kWarpsPerCta - 1) / kWarpsPerCta;
    inv_fft1024_kernel<<<ctas, kWarpsPerCta * 32, 0, stream>>>(plan.jobs.device(), njobs, plan.entries.device(),
        reinterpret_cast<const float2*>(partials.ptr), red, tw.ptr, fade_out.ptr, fade_in.ptr);
  }
  else
  {
    int max_entries = 1;
    for (auto& j : plan.jobs.host) max_entries = std::max(max_entries, j.n_entries);
    launch_inv_kernel(inv_groups(max_entries), njobs, plan, red, ga, nullptr, hc);
  }
  SSR_CUDA(cudaGetLastError());
  count_launch();
  if (cur) cur->ifft_launches++;
}

bool Engine::is_pinned(const void* p, size_t bytes)
{
  for (auto& r : pinned_cache)
    if (r.p == p && r.bytes >= bytes) return r.pinned;
  auto host_pinned = [](const void* q)
  {
    cudaPointerAttributes attr;
    if (cudaPointerGetAttributes(&attr, q) != cudaSuccess) { cudaGetLastError(); return false; }
    return attr.type == cudaMemoryTypeHost;
  };
  const bool pinned = bytes > 0 && host_pinned(p)
      && host_pinned(static_cast<const char*>(p) + bytes - 1);
  if (pinned_cache.size() >= 64) pinned_cache.erase(pinned_cache.begin());
  pinned_cache.push_back({p, bytes, pinned});
  return pinned;
}

void Engine::upload_weights(const float* w, size_t n)
{
  if (n == 0) return;
  // steady state (weights unchanged since the last upload): nothing to do
  if (wtab.uploaded == n && wtab.host.size() == n
      && std::memcmp(wtab.host.data(), w, n * sizeof(float)) == 0) return;
  wtab.host.assign(w, w + n);
  wtab.upload(stream);
}

void Engine::stage_begin()
{
  if (!timing || (timing_counter++ % uint64_t(timing_every)) != 0) { cur = nullptr; return; }
  StageEvents se{};
  if (!free_events.empty()) { se = free_events.back(); free_events.pop_back(); }
  else for (auto& ev : se.ev) SSR_CUDA(cudaEventCreate(&ev));
  se.mac_bytes = 0; se.mac_terms = 0; se.mac_launches = se.fft_launches = se.ifft_launches = 0;
  pending.push_back(se);
  cur = &pending.back();
  SSR_CUDA(cudaEventRecord(cur->ev[0], stream));
}

void Engine::stage_mark(int i) { if (cur) SSR_CUDA(cudaEventRecord(cur->ev[i], stream)); }

void Engine::stage_end()
{
  cur = nullptr;
  if (pending.size() > 4096) fold_stats();
}

void Engine::fold_stats()
{
  if (pending.empty()) return;
  sync();
  for (auto& p : pending)
  {
    float a = 0, b = 0, c = 0, d = 0;
    cudaEventElapsedTime(&a, p.ev[0], p.ev[1]);
    cudaEventElapsedTime(&b, p.ev[1], p.ev[2]);
    cudaEventElapsedTime(&c, p.ev[2], p.ev[3]);
    cudaEventElapsedTime(&d, p.ev[0], p.ev[3]);
    stats.ms_fft += a; stats.ms_mac += b; stats.ms_ifft += c; stats.ms_total += d;
    stats.blocks++;
    stats.mac_launches += p.mac_launches; stats.fft_launches += p.fft_launches;
    stats.ifft_launches += p.ifft_launches;
    stats.mac_bytes += p.mac_bytes; stats.mac_terms += p.mac_terms;
    free_events.push_back(p);
  }
  pending.clear();
  // partitions read by generated-terms MAC launches (counted on the device)
  unsigned long long hx[2] = {0, 0};
  SSR_CUDA(cudaMemcpy(hx, d_dyn_stats.ptr, sizeof(hx), cudaMemcpyDeviceToHost));
  stats.mac_bytes += uint64_t(8) * B * ((hx[0] - dyn_h_seen) + (hx[1] - dyn_x_seen));
  stats.mac_terms += hx[0] - dyn_h_seen;
  dyn_h_seen = hx[0]; dyn_x_seen = hx[1];
}

// ------------------------------------------------------------------ FilterSet
FilterSet::FilterSet(Engine* e_, int channels_, int partitions_)
  : e(e_), channels(channels_), partitions(partitions_)
{
  if (channels <= 0 || partitions <= 0) throw Error(SSR_ERR_INVALID, "ssr_b200: empty filter set");
  e->use_device();
  data.resize(size_t(channels) * partitions * e->B);
  valid.assign(channels, 0);
  flags.assign(size_t(channels) * partitions, 0);
}

void FilterSet::touch_flags()
{
  flags_dirty = true;
  ++e->flags_epoch;
}

const unsigned char* FilterSet::device_flags()
{
  if (flags_dirty || !d_flags.ptr)
  {
    e->use_device();
    if (d_flags.count < flags.size()) { e->sync(); d_flags.resize(flags.size()); }
    // pageable source: the driver stages it before returning; ordered on the stream
    // behind the kernels that still read the previous flags
    SSR_CUDA(cudaMemcpyAsync(d_flags.ptr, flags.data(), flags.size(), cudaMemcpyHostToDevice, e->stream));
    all_set.assign(size_t(channels), 1);
    for (int ch = 0; ch < channels; ++ch)
      for (int p = 0; p < partitions; ++p)
        if (!flags[size_t(ch) * partitions + p]) { all_set[ch] = 0; break; }
    flags_dirty = false;
  }
  return d_flags.ptr;
}

const unsigned char* FilterSet::device_flags_row(int ch)
{
  const unsigned char* base = device_flags();
  return all_set[ch] ? nullptr : base + size_t(ch) * partitions;
}

static void run_fwd_jobs(Engine* e, std::vector<FwdJob>& jobs)
{
  // chunked so the job table stays small
  const size_t chunk = 1 << 16;
  DeviceArray<FwdJob> d(std::min(chunk, jobs.size()));
  for (size_t off = 0; off < jobs.size(); off += chunk)
  {
    const size_t n = std::min(chunk, jobs.size() - off);
    SSR_CUDA(cudaMemcpyAsync(d.ptr, jobs.data() + off, n * sizeof(FwdJob),
          cudaMemcpyHostToDevice, e->stream));
    e->launch_fwd(d.ptr, int(n));
    e->sync();
  }
}

void FilterSet::load_time(int first_channel, int nch, const float* host, long frames,
                          long frame_stride, long channel_stride)
{
  if (first_channel < 0 || nch <= 0 || first_channel + nch > channels || frames < 0 || !host)
    throw Error(SSR_ERR_INVALID, "ssr_b200: filterset_load_time: bad arguments");
  e->use_device();
  const int B = e->B;
  // extent of the host array touched
  const long extent = (frames > 0)
    ? (frames - 1) * frame_stride + (nch - 1) * channel_stride + 1 : 0;
  DeviceArray<float> d_time{static_cast<size_t>(std::max<long>(extent, 1))};
  if (extent > 0)
    SSR_CUDA(cudaMemcpyAsync(d_time.ptr, host, size_t(extent) * sizeof(float),
          cudaMemcpyHostToDevice, e->stream));
  const int nvalid = int(std::min<long>(partitions, (frames + B - 1) / B));
  for (int c = 0; c < nch; ++c)
  {
    const int ch = first_channel + c;
    valid[ch] = nvalid;
    for (int p = 0; p < partitions; ++p) flags[size_t(ch) * partitions + p] = p < nvalid;
    touch_flags();
  }
  FwdBatch batch{};
  batch.time = d_time.ptr;
  batch.frame_stride = frame_stride; batch.channel_stride = channel_stride;
  batch.frames = frames;
  batch.set = data.ptr; batch.partitions = partitions; batch.first_channel = first_channel;
  batch.synth = 0;
  e->launch_fwd_batch(batch, nvalid, nch);
  e->sync();  // d_time goes out of scope
}

void FilterSet::set_partition_time(int ch, int p, const float* taps, int n)
{
  if (ch < 0 || ch >= channels || p < 0 || p >= partitions || n < 0 || n > e->B)
    throw Error(SSR_ERR_INVALID, "ssr_b200: set_partition_time: bad arguments");
  e->use_device();
  if (n == 0)
  {
    // chunk == 0 -> zero flag, no FFT (convolver.h:218-221)
    flags[size_t(ch) * partitions + p] = 0;
    touch_flags();
    return;
  }
  DeviceArray<float> d{size_t(n)};
  SSR_CUDA(cudaMemcpyAsync(d.ptr, taps, size_t(n) * sizeof(float), cudaMemcpyHostToDevice, e->stream));
  std::vector<FwdJob> jobs(1);
  jobs[0] = FwdJob{};
  jobs[0].first.ptr = d.ptr; jobs[0].first.stride = 1; jobs[0].first.count = n;
  jobs[0].second.stride = 1;
  jobs[0].dst = part(ch, p);
  run_fwd_jobs(e, jobs);
  flags[size_t(ch) * partitions + p] = 1;
  touch_flags();
  valid[ch] = std::max(valid[ch], p + 1);
}

void FilterSet::generate(int first_channel, int nch, long taps, uint64_t seed, uint64_t id_offset,
                         const float* envelope, float scale, uint64_t tap_offset)
{
  if (first_channel < 0 || nch <= 0 || first_channel + nch > channels || taps <= 0)
    throw Error(SSR_ERR_INVALID, "ssr_b200: filterset_generate: bad arguments");
  e->use_device();
  const int B = e->B;
  DeviceArray<float> d_env;
  if (envelope)
  {
    d_env.resize(size_t(taps));
    SSR_CUDA(cudaMemcpyAsync(d_env.ptr, envelope, size_t(taps) * sizeof(float),
          cudaMemcpyHostToDevice, e->stream));
  }
  const int nvalid = int(std::min<long>(partitions, (taps + B - 1) / B));
  for (int c = 0; c < nch; ++c)
  {
    const int ch = first_channel + c;
    valid[ch] = nvalid;
    for (int p = 0; p < partitions; ++p) flags[size_t(ch) * partitions + p] = p < nvalid;
    touch_flags();
  }
  FwdBatch batch{};
  // (tap t of the slice is tap tap_offset + t of the generated impulse response; the envelope table
  // passed in covers the slice, the kernels index it by the absolute tap)
  batch.time = d_env.ptr ? d_env.ptr - tap_offset : nullptr;
  batch.tap_offset = tap_offset;
  batch.frame_stride = 1; batch.channel_stride = 0;
  batch.frames = taps;
  batch.set = data.ptr; batch.partitions = partitions; batch.first_channel = first_channel;
  batch.synth = 1; batch.scale = scale; batch.seed = seed; batch.id_offset = id_offset;
  e->launch_fwd_batch(batch, nvalid, nch);
  e->sync();  // d_env goes out of scope
}

void FilterSet::lerp_plan(int dch, const FilterSet& a, int ach, const FilterSet& b, int bch, float t,
                          std::vector<LerpJob>& jobs)
{
  if (dch < 0 || dch >= channels || ach < 0 || ach >= a.channels || bch < 0 || bch >= b.channels)
    throw Error(SSR_ERR_INVALID, "ssr_b200: filterset_lerp: bad channel");
  bool changed = false;
  for (int p = 0; p < partitions; ++p)
  {
    const float2* pa = a.part_or_null(ach, p);
    const float2* pb = b.part_or_null(bch, p);
    const uint8_t flag = (pa || pb) ? 1 : 0;
    uint8_t& cur = flags[size_t(dch) * partitions + p];
    changed |= (cur != flag);
    cur = flag;
    if (!flag) continue;
    jobs.push_back({reinterpret_cast<const float4*>(pa), reinterpret_cast<const float4*>(pb),
        reinterpret_cast<float4*>(part(dch, p)), 1.0f - t, t});
  }
  if (changed) touch_flags();
}

void FilterSet::lerp_from(int dch, const FilterSet& a, int ach, const FilterSet& b, int bch, float t)
{
  e->use_device();
  std::vector<LerpJob> jobs;
  lerp_plan(dch, a, ach, b, bch, t, jobs);
  if (jobs.empty()) return;
  DeviceArray<LerpJob> d(jobs.size());
  SSR_CUDA(cudaMemcpyAsync(d.ptr, jobs.data(), jobs.size() * sizeof(LerpJob),
        cudaMemcpyHostToDevice, e->stream));
  lerp_kernel<<<int(jobs.size()), kThreads, 0, e->stream>>>(d.ptr, e->B / 2);
  SSR_CUDA(cudaGetLastError());
  e->sync();  // `jobs` / `d` go out of scope
}

// packed complex -> the reference's sorted halfcomplex layout (convolver.h:236-268):
// group g of 8 floats = [Re X_{4g..4g+3} | Im X_{4g..4g+3}], slot 4 of group 0 = Re X_B
static void packed_to_sorted(const float2* packed, int B, float* sorted)
{
  for (int k = 0; k < B; ++k)
  {
    const int g = k >> 2, j = k & 3;
    sorted[8 * g + j] = packed[k].x;
    sorted[8 * g + 4 + j] = packed[k].y;
  }
}

// the inverse mapping (filterset_upload)
static void sorted_to_packed(const float* sorted, int B, float2* packed)
{
  for (int k = 0; k < B; ++k)
  {
    const int g = k >> 2, j = k & 3;
    packed[k] = make_float2(sorted[8 * g + j], sorted[8 * g + 4 + j]);
  }
}

// writes one partition spectrum given in the reference's layout (transform_nested
// with a host functor, parity tooling); zero != 0 sets the zero flag instead
void FilterSet::upload(int ch, int p, const float* sorted_in, int zero)
{
  if (ch < 0 || ch >= channels || p < 0 || p >= partitions || (!zero && !sorted_in))
    throw Error(SSR_ERR_INVALID, "ssr_b200: filterset_upload: bad arguments");
  e->use_device();
  touch_flags();
  if (zero) { flags[size_t(ch) * partitions + p] = 0; }
  else
  {
    const int B = e->B;
    std::vector<float2> tmp(B);
    sorted_to_packed(sorted_in, B, tmp.data());
    e->sync();  // kernels in flight may still read the partition
    SSR_CUDA(cudaMemcpy(part(ch, p), tmp.data(), B * sizeof(float2), cudaMemcpyHostToDevice));
    flags[size_t(ch) * partitions + p] = 1;
  }
  int v = 0;
  for (int q = 0; q < partitions; ++q) if (flags[size_t(ch) * partitions + q]) v = q + 1;
  valid[ch] = v;
}

void FilterSet::download(int ch, int p, float* sorted_out, int* zero)
{
  if (ch < 0 || ch >= channels || p < 0 || p >= partitions)
    throw Error(SSR_ERR_INVALID, "ssr_b200: filterset_download: bad index");
  e->use_device();
  const int B = e->B;
  const bool z = !flags[size_t(ch) * partitions + p];
  if (zero) *zero = z ? 1 : 0;
  if (z) { std::fill(sorted_out, sorted_out + 2 * B, 0.0f); return; }
  std::vector<float2> tmp(B);
  e->sync();
  SSR_CUDA(cudaMemcpy(tmp.data(), part(ch, p), B * sizeof(float2), cudaMemcpyDeviceToHost));
  packed_to_sorted(tmp.data(), B, sorted_out);
}

// ------------------------------------------------------------------ FilterPtrTable
FilterPtrTable::FilterPtrTable(int partitions)
  : P(partitions)
{}

void FilterPtrTable::set_static(const FilterSet& fs, int ch)
{
  // fewer partitions given -> the rest is empty; more -> ignored (convolver.h:733-738)
  base = Event{0, fs.part(ch, 0), size_t(fs.e->B), fs.flags.data() + size_t(ch) * fs.partitions,
               fs.partitions, nullptr};
  events.clear();
  ++version;
}

void FilterPtrTable::set_static_ptrs(const std::vector<const float2*>& parts)
{
  base = Event{0, nullptr, 0, nullptr, 0, std::make_shared<std::vector<const float2*>>(parts)};
  events.clear();
  ++version;
}

void FilterPtrTable::push_event(Event ev)
{
  // a second set_filter() before the next rotate overwrites the same queue slots
  if (!events.empty() && events.back().t_set == ev.t_set) events.back() = std::move(ev);
  else events.push_back(std::move(ev));
  ++version;
}

void FilterPtrTable::set_filter(const FilterSet& fs, int ch)
{
  Event ev{T, fs.part(ch, 0), size_t(fs.e->B), fs.flags.data() + size_t(ch) * fs.partitions,
           fs.partitions, nullptr};
  push_event(std::move(ev));
}

void FilterPtrTable::set_filter_ptrs(const std::vector<const float2*>& parts)
{
  Event ev{T, nullptr, 0, nullptr, 0, std::make_shared<std::vector<const float2*>>(parts)};
  push_event(std::move(ev));
}

bool FilterPtrTable::queues_empty() const
{
  // "some non-null pointer in the last queue" (convolver.h:669-676): an event
  // whose partition P-1 is not active yet
  if (P < 2) return true;
  return events.empty() || events.back().t_set + uint64_t(P - 1) <= T;
}

void FilterPtrTable::rotate_queues()
{
  ++T;
  // an event all of whose partitions are active supersedes everything older:
  // fold the newest such event into the base table
  int fold = -1;
  for (int i = int(events.size()) - 1; i >= 0; --i)
    if (events[i].t_set + uint64_t(P - 1) <= T) { fold = i; break; }
  if (fold >= 0)
  {
    // partitions for which a newer event already qualifies are taken from it in current()
    base = events[fold];
    events.erase(events.begin(), events.begin() + fold + 1);
  }
  ++version;
}

const std::vector<const float2*>& FilterPtrTable::current()
{
  if (_ptrs_version == version) return _ptrs;
  _ptrs.resize(size_t(P));
  int idx = int(events.size()) - 1;
  for (int p = 0; p < P; ++p)
  {
    // t_set + p grows with p, so the newest qualifying event only moves backwards
    while (idx >= 0 && events[idx].t_set + uint64_t(p) > T) --idx;
    _ptrs[p] = idx >= 0 ? events[idx].part(p) : base.part(p);
  }
  _ptrs_version = version;
  return _ptrs;
}

bool FilterPtrTable::references(const float2* lo, const float2* hi) const
{
  for (int p = 0; p < P; ++p) { const float2* q = base.part(p); if (q >= lo && q < hi) return true; }
  for (auto& ev : events)
    for (int p = 0; p < P; ++p) { const float2* q = ev.part(p); if (q >= lo && q < hi) return true; }
  return false;
}

// ------------------------------------------------------------------ Input
Input::Input(Engine* e_, int partitions)
  : e(e_), id(-1), P(partitions), slots(partitions + kTimeBatch - 1)
{
  if (partitions <= 0) throw Error(SSR_ERR_INVALID, "ssr_b200: Input needs partitions > 0");
  e->use_device();
  const int B = e->B;
  ring.resize(size_t(slots) * B);
  prev.resize(B); cur.resize(B);
  stage.resize(B);
  SSR_CUDA(cudaMemsetAsync(ring.ptr, 0, size_t(slots) * B * sizeof(float2), e->stream));
  SSR_CUDA(cudaMemsetAsync(prev.ptr, 0, B * sizeof(float), e->stream));
  id = e->register_input(this);
  job.host.assign(1, make_job(cur.ptr));
  job.upload(e->stream);
  e->sync();
}

Input::~Input() { e->unregister_input(id); }

FwdJob Input::make_job(const float* d_cur)
{
  FwdJob j{};
  j.first.ptr = prev.ptr; j.first.stride = 1; j.first.count = e->B;
  j.second.ptr = d_cur; j.second.stride = 1; j.second.count = e->B;
  j.dst = nullptr;
  j.ring = ring.ptr;
  j.head = e->d_heads.ptr + id;
  j.parts = slots;
  j.save_second = prev.ptr;
  j.level = nullptr;
  return j;
}

void Input::add_block(const float* x)
{
  e->use_device();
  // a block nobody convolved yet is transformed now (the delay line advances once
  // per add_block); then wait until the staging buffer is free
  if (fwd_pending) flush();
  // (a fused launch has been waited for by convolve(): the buffer is free then)
  if (stage_busy) { e->sync(); stage_busy = false; }
  std::memcpy(stage.ptr, x, e->B * sizeof(float));
  e->stats.h2d_bytes += e->B * sizeof(float);
  fwd_pending = true;
  if (!e->level1_fused) flush();
}

void Input::flush()
{
  if (!fwd_pending) return;
  e->use_device();
  SSR_CUDA(cudaMemcpyAsync(cur.ptr, stage.ptr, e->B * sizeof(float), cudaMemcpyHostToDevice, e->stream));
  e->launch_fwd(job.device(), 1);
  fwd_pending = false;
  stage_busy = true;   // the copy out of the staging buffer is asynchronous
}

void Input::download(int p, float* sorted_out)
{
  if (p < 0 || p >= P) throw Error(SSR_ERR_INVALID, "ssr_b200: input_download: bad partition");
  e->use_device();
  flush();
  e->sync();
  int head = 0;
  SSR_CUDA(cudaMemcpy(&head, e->d_heads.ptr + id, sizeof(int), cudaMemcpyDeviceToHost));
  int slot = head - p; if (slot < 0) slot += slots;
  std::vector<float2> tmp(e->B);
  SSR_CUDA(cudaMemcpy(tmp.data(), ring.ptr + size_t(slot) * e->B, e->B * sizeof(float2), cudaMemcpyDeviceToHost));
  packed_to_sorted(tmp.data(), e->B, sorted_out);
}

// ------------------------------------------------------------------ Output
Output::Output(Input* in_, ssr_output_kind kind_)
  : in(in_), e(in_->e), kind(kind_), table(in_->P)
{
  e->use_device();
  d_res.resize(e->B); h_res.resize(e->B);
  h_flag.resize(16);
  h_flag.ptr[0] = 0;
  mac.MT = 1;
}

const float* Output::convolve(float weight)
{
  e->use_device();
  const int B = e->B;
  // The MAC / inverse tables only depend on the pointer table: rebuild them when
  // set_filter / rotate_queues / bind changed it, otherwise relaunch as they are.
  const bool zero = (weight == 0.0f);
  if (planned_flags_epoch != e->flags_epoch)
  {
    // a filter of this engine was edited in place (prepare_partition, upload, transform_nested
    // into a filter that is bound or queued here): zero flags are read live, like the
    // reference's _multiply_spectra does
    ++table.version;
    planned_flags_epoch = e->flags_epoch;
  }
  if (planned_version != table.version || planned_zero != zero)
  {
    mac.clear();
    mac.begin_group();
    const auto& ptrs = table.current();
    for (int p = 0; p < table.P; ++p)
    {
      // zero-flagged partitions are skipped (convolver.h:561)
      if (!ptrs[p]) continue;
      const float4* h = reinterpret_cast<const float4*>(ptrs[p]);
      mac.add_term(in->id, p, 0, &h);
    }
    mac.end_group();
    planned_terms = int(mac.meta.host.size());
    inv.clear();
    if (planned_terms == 0 || zero)
    {
      // nothing contributed: the reference returns its zero-filled buffer
      // without an IFFT (convolver.h:448-452)
      inv.jobs.host.push_back({0, 0, d_res.ptr, nullptr});
    }
    else
    {
      mac.make_work(e->mac_slots(mac.MT), e->split_cost(mac.MT), e->mac_waves);
      mac.upload(e->stream);
      e->ensure_partials(size_t(mac.npartials) * mac.MT);
      inv.entries.host.push_back({0, mac.groups[0].nsplit, 1, 0, 0, 1.0f / float(2 * B)});
      inv.jobs.host.push_back({0, 1, d_res.ptr, nullptr});
    }
    inv.upload(e->stream);
    planned_version = table.version;
    planned_zero = zero;
  }
  if (e->level1_fused && !zero && planned_terms > 0 && planned_terms <= kFusedMaxTerms)
    return convolve_fused(weight);
  in->flush();
  if (planned_terms > 0 && !zero)
  {
    // other objects of this engine may have grown the shared partial buffer
    e->ensure_partials(size_t(mac.npartials) * mac.MT);
    e->upload_weights(&weight, 1);
    e->launch_mac(mac, 0, 0);
  }
  e->launch_inv(inv);
  SSR_CUDA(cudaMemcpyAsync(h_res.ptr, d_res.ptr, B * sizeof(float), cudaMemcpyDeviceToHost, e->stream));
  e->stats.d2h_bytes += B * sizeof(float);
  e->sync();
  return h_res.ptr;
}

// One launch, no copy engine, no stream synchronisation (see level1_fused_kernel).
const float* Output::convolve_fused(float weight)
{
  const int B = e->B;
  FusedParams p{};
  p.fwd = in->make_job(in->stage.ptr);   // the kernel reads the block from mapped host memory
  p.do_fwd = in->fwd_pending ? 1 : 0;
  p.term_meta = mac.meta.device();
  p.term_h = mac.hptr.device();
  p.nterms = planned_terms;
  p.xring = e->d_xring.ptr;
  p.xparts = e->d_xparts.ptr;
  p.heads = e->d_heads.ptr;
  p.scale = weight / float(2 * B);       // weight / n (convolver.h:458)
  p.out = h_res.ptr;
  p.flag = h_flag.ptr;
  p.seq = ++fused_seq;
  p.B = B;
  e->launch_level1_fused(p);
  in->fwd_pending = false;
  e->stats.d2h_bytes += B * sizeof(float);
  // spin on the completion flag; a failed launch / faulted context ends the wait
  volatile int* flag = h_flag.ptr;
  for (unsigned spins = 0; *flag != p.seq; ++spins)
  {
    if ((spins & 0xfffu) == 0xfffu)
    {
      const cudaError_t q = cudaStreamQuery(e->stream);
      if (q != cudaSuccess && q != cudaErrorNotReady)
        throw Error(SSR_ERR_CUDA, std::string("level1_fused_kernel: ") + cudaGetErrorString(q));
      if (q == cudaSuccess && *flag != p.seq)
        throw Error(SSR_ERR_CUDA, "level1_fused_kernel finished without delivering its block");
    }
  }
  return h_res.ptr;
}

}  // namespace ssr_b200
