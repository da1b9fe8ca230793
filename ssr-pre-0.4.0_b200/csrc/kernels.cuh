// kernels.cuh -- hand-written sm_100a kernels of the partitioned-convolution path.
// This header holds the shared helpers, the FFT core and the forward kernels and
// includes mac.cuh (MAC kernels), inverse.cuh (partial reduction, inverse transform +
// crossfade, fused output gather) and level1.cuh (fused Level-1 launch);
// fft1024.cuh has the warp-per-transform FFTs.
//
//   fwd_fft_kernel   batched real->packed-complex FFT of [first half | second half]
//                    (input blocks: [prev | cur]; filter partitions: [taps | 0]).
//                    Replaces FFTW R2HC + _sort_coefficients
//                    (apf/apf/convolver.h:144-148, 176-181, 209-268, 324-375).
//   mac_tma_kernel / mac_kernel
//                    frequency-domain delay-line complex multiply-accumulate
//                    acc[row, k] = sum_terms w * X[n, p, k] * H[row, n, p, k].
//                    Replaces _multiply_spectra / _multiply_partition_simd
//                    (convolver.h:503-572).  HBM-bound, FP32 FMA, no tensor cores.
//   inv_fft_kernel   sum of split partial accumulators + packed-complex->real
//                    inverse FFT + 1/(2B) scale + keep-second-half + raised-cosine
//                    crossfade + accumulate into the output block.  Replaces
//                    _unsort_coefficients + FFTW HC2R + scaling
//                    (convolver.h:448-463, 574-613) and
//                    CombineChannelsCrossfadeCopy (apf/apf/combine_channels.h:
//                    297-335, 366-403).
//
// Spectrum layout ("packed complex"): a partition spectrum is B float2;
// element k (1 <= k < B) is bin k, element 0 is (Re X_0, Re X_B) -- DC and
// Nyquist are both real.  The reference's "sorted halfcomplex" layout
// (convolver.h:236-268) is only produced on download for parity checks.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace ssrk
{

constexpr int kThreads = 256;

// ---------------------------------------------------------------- synthetic data
// counter-based generator shared by host and device (bit-identical)
__host__ __device__ inline uint64_t mix64(uint64_t z)
{
  z ^= z >> 30; z *= 0xBF58476D1CE4E5B9ull;
  z ^= z >> 27; z *= 0x94D049BB133111EBull;
  z ^= z >> 31;
  return z;
}

__host__ __device__ inline uint64_t synth_key(uint64_t seed, uint64_t stream)
{
  return mix64(seed ^ mix64(stream + 0x9E3779B97F4A7C15ull));
}

// uniform in [-1, 1) on a 2^-23 grid: exact in float32 on host and device
__host__ __device__ inline float synth_uniform(uint64_t key, uint64_t index)
{
  const uint64_t h = mix64(key + index * 0x9E3779B97F4A7C15ull);
  return float(uint32_t(h >> 40)) * (1.0f / 8388608.0f) - 1.0f;
}

// ---------------------------------------------------------------- small helpers
// Programmatic dependent launch: a kernel launched with the programmatic-stream-
// serialization attribute may be scheduled while its predecessor in the stream is
// still draining; it must not touch the predecessor's results before this wait
// (which also makes them visible).  Without the attribute this is a no-op.
__device__ __forceinline__ void grid_dependency_wait()
{
  asm volatile("griddepcontrol.wait;" ::: "memory");
}

// Lets the next kernel of the stream (launched with the programmatic attribute) start
// once every CTA of this grid has got here or exited; its own griddepcontrol.wait
// still waits for this grid's completion.
__device__ __forceinline__ void grid_launch_dependents()
{
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}

__device__ __forceinline__ float2 cmul(float2 a, float2 b)
{
  return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}

__device__ __forceinline__ float4 ldg_stream(const float4* p)
{
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0, %1, %2, %3}, [%4];"
      : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
  return r;
}

// L2 eviction-priority hints: the H stream (read once per block, 25 GB >> L2)
// is marked evict-first, the frequency-domain delay line X (re-read by every
// output group) evict-last, so H does not push X out of the 126 MB L2.
__device__ __forceinline__ uint64_t l2_policy_evict_first()
{
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}

__device__ __forceinline__ uint64_t l2_policy_evict_last()
{
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}

__device__ __forceinline__ uint64_t l2_policy_evict_normal()
{
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_normal.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}

__device__ __forceinline__ float4 ldg_stream_hint(const float4* p, uint64_t pol)
{
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.f32 {%0, %1, %2, %3}, [%4], %5;"
      : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p), "l"(pol));
  return r;
}

// max over the CTA (all threads must call); result valid in thread 0
__device__ __forceinline__ float block_max(float v)
{
  __shared__ float s_max[kThreads / 32];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  if ((threadIdx.x & 31) == 0) s_max[threadIdx.x >> 5] = v;
  __syncthreads();
  if (threadIdx.x < 32)
  {
    v = threadIdx.x < kThreads / 32 ? s_max[threadIdx.x] : 0.0f;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  }
  return v;
}

// ---------------------------------------------------------------- FFT core
// Complex FFT of size m in shared memory, Stockham autosort: radix-4 stages, one radix-2
// stage when the power of two in m is odd, then -- block sizes that are multiples of 8
// but not powers of two (the reference accepts every B % 8 == 0, convolver.h:163-166) --
// one generic radix-r stage per remaining odd prime factor r (direct r-point DFT per
// butterfly: cheap for 3, 5, 7; correct for any r).  `a` holds the
// input; the result ends up in the returned buffer (a or b).  tw[i] =
// exp(-2 pi i * i / (2m)), i in [0, 2m); INVERSE conjugates the twiddles.
// barrier `bar` over kThreads threads (0 = the whole CTA of a kThreads-thread kernel;
// the inverse kernel runs one kThreads-thread group per accumulator entry, each on
// its own named barrier)
__device__ __forceinline__ void group_sync(int bar)
{
  asm volatile("bar.sync %0, %1;" :: "r"(bar), "n"(kThreads) : "memory");
}

// TWS: `tw` is a shared-memory copy of the twiddle table (the latency-bound
// block-loop kernels stage it once instead of going to L1/L2 in every stage)
template<bool INVERSE, bool TWS = false>
__device__ __forceinline__ float2* fft_smem(float2* a, float2* b, int m,
                                            const float2* __restrict__ tw,
                                            const int t = threadIdx.x, const int bar = 0)
{
  float2* src = a;
  float2* dst = b;
  int Ns = 1;
  int rem = m;
  while (rem > 1)
  {
    if ((rem & 3) == 0)
    {
      const int q = m >> 2;
      // twiddle index step: e^{-2 pi i k / (4 Ns)} = tw[k * (2m / (4 Ns))]
      const int tstep = (2 * m) / (4 * Ns);
      for (int j = t; j < q; j += kThreads)
      {
        const int k = j & (Ns - 1);
        float2 v0 = src[j];
        float2 v1 = src[j + q];
        float2 v2 = src[j + 2 * q];
        float2 v3 = src[j + 3 * q];
        if (Ns > 1)
        {
          float2 w1 = TWS ? tw[k * tstep] : __ldg(&tw[k * tstep]);
          float2 w2 = TWS ? tw[2 * k * tstep] : __ldg(&tw[2 * k * tstep]);
          float2 w3 = TWS ? tw[3 * k * tstep] : __ldg(&tw[3 * k * tstep]);
          if (INVERSE) { w1.y = -w1.y; w2.y = -w2.y; w3.y = -w3.y; }
          v1 = cmul(v1, w1); v2 = cmul(v2, w2); v3 = cmul(v3, w3);
        }
        const float2 s02 = make_float2(v0.x + v2.x, v0.y + v2.y);
        const float2 d02 = make_float2(v0.x - v2.x, v0.y - v2.y);
        const float2 s13 = make_float2(v1.x + v3.x, v1.y + v3.y);
        const float2 d13 = make_float2(v1.x - v3.x, v1.y - v3.y);
        // forward: -i * d13 = (d13.y, -d13.x); inverse: +i * d13 = (-d13.y, d13.x)
        const float2 r = INVERSE ? make_float2(-d13.y, d13.x) : make_float2(d13.y, -d13.x);
        const int j0 = ((j - k) << 2) + k;
        dst[j0]          = make_float2(s02.x + s13.x, s02.y + s13.y);
        dst[j0 + Ns]     = make_float2(d02.x + r.x, d02.y + r.y);
        dst[j0 + 2 * Ns] = make_float2(s02.x - s13.x, s02.y - s13.y);
        dst[j0 + 3 * Ns] = make_float2(d02.x - r.x, d02.y - r.y);
      }
      Ns <<= 2; rem >>= 2;
    }
    else if ((rem & 1) == 0)
    {
      const int q = m >> 1;
      const int tstep = (2 * m) / (2 * Ns);
      for (int j = t; j < q; j += kThreads)
      {
        const int k = j & (Ns - 1);
        float2 v0 = src[j];
        float2 v1 = src[j + q];
        if (Ns > 1)
        {
          float2 w1 = TWS ? tw[k * tstep] : __ldg(&tw[k * tstep]);
          if (INVERSE) w1.y = -w1.y;
          v1 = cmul(v1, w1);
        }
        const int j0 = ((j - k) << 1) + k;
        dst[j0]      = make_float2(v0.x + v1.x, v0.y + v1.y);
        dst[j0 + Ns] = make_float2(v0.x - v1.x, v0.y - v1.y);
      }
      Ns <<= 1; rem >>= 1;
    }
    else
    {
      // smallest odd factor r of what is left (the powers of two are gone, so Ns may
      // stop being a power of two from here on: % instead of &)
      int r = 3;
      while (rem % r) r += 2;
      const int q = m / r;
      const int tstep = (2 * m) / (r * Ns);   // e^{-2 pi i k / (r Ns)} = tw[k * tstep]
      const int rstep = (2 * m) / r;          // e^{-2 pi i / r}        = tw[rstep]
      for (int j = t; j < q; j += kThreads)
      {
        const int k = j % Ns;
        const int j0 = (j - k) * r + k;
        for (int pp = 0; pp < r; ++pp)
        {
          float2 acc = make_float2(0.f, 0.f);
          for (int qq = 0; qq < r; ++qq)
          {
            int idx = qq * k * tstep + ((pp * qq) % r) * rstep;   // both exponents on the 2m-entry table
            if (idx >= 2 * m) idx -= 2 * m;
            float2 w = TWS ? tw[idx] : __ldg(&tw[idx]);
            if (INVERSE) w.y = -w.y;
            const float2 v = cmul(src[j + qq * q], w);
            acc.x += v.x; acc.y += v.y;
          }
          dst[j0 + pp * Ns] = acc;
        }
      }
      Ns *= r; rem /= r;
    }
    group_sync(bar);
    float2* tmp = src; src = dst; dst = tmp;
  }
  return src;
}

// copies the 2B-entry twiddle table into shared memory (all `nthreads` threads of
// the CTA take part; the caller's next CTA-wide barrier publishes it)
__device__ __forceinline__ void stage_twiddles(float2* stw, const float2* __restrict__ tw, int B,
                                               int tid, int nthreads)
{
  const float4* src = reinterpret_cast<const float4*>(tw);
  float4* dst = reinterpret_cast<float4*>(stw);
  for (int i = tid; i < B; i += nthreads) dst[i] = __ldg(src + i);
}

// ---------------------------------------------------------------- forward
struct FwdHalf
{
  const float* ptr;   // memory source (or envelope table when synth != 0), may be null = zeros
  long stride;        // element stride (de-interleaving libsndfile frames x channels)
  int count;          // valid samples (<= B); the rest is zero padding
  int synth;          // 1: generate u(key, first_index + i) * ptr[env_offset + i] * scale
  uint64_t key;
  uint64_t first_index;
  float scale;
  int pad;
};

struct FwdJob
{
  FwdHalf first, second;
  float2* dst;        // explicit destination spectrum, or null -> ring mode
  float2* ring;       // ring base (parts x B float2)
  int* head;          // ring head (device); advanced by this kernel
  int parts;          // ring length
  float* save_second; // if non-null: B floats <- second half (becomes "prev")
  float* level;       // if non-null: max |second half| (level meter,
                      //  src/rendererbase.h:431-435 math::max_amplitude)
};

__device__ __forceinline__ float fwd_sample(const FwdHalf& h, int i)
{
  if (i >= h.count) return 0.0f;
  if (h.synth)
  {
    const float u = synth_uniform(h.key, h.first_index + uint64_t(i));
    const float env = h.ptr ? __ldg(h.ptr + h.first_index + i) : 1.0f;
    return (u * env) * h.scale;
  }
  return h.ptr ? h.ptr[long(i) * h.stride] : 0.0f;
}

// One CTA per transform.  B = block size = complex FFT size; smem = 2 * B float2
// (TWS: + 2 * B float2 for the staged twiddle table).
// DEFER_HEAD: the new ring head is NOT published (the block's inverse kernel commits it,
// HeadCommit) so that a MAC launch running concurrently sees a stable value.
template<bool TWS = false, bool DEFER_HEAD = false>
__device__ __forceinline__ void fwd_fft_body(const FwdJob& job, int B, const float2* __restrict__ tw_global)
{
  extern __shared__ float2 smem[];
  float2* a = smem;
  float2* b = smem + B;
  const int t = threadIdx.x;
  const int half = B >> 1;
  const float2* tw = tw_global;
  if (TWS)
  {
    stage_twiddles(smem + 2 * B, tw_global, B, t, kThreads);  // published by the barrier after the input load
    tw = smem + 2 * B;
  }

  int new_head = 0;
  if (!job.dst)
  {
    // (volatile: with Engine::block_overlap the head was committed by a kernel that may still be running)
    new_head = *reinterpret_cast<volatile int*>(job.head) + 1;
    if (new_head >= job.parts) new_head = 0;
  }

  // z[j] = x[2j] + i x[2j+1], x = [first | second]
  for (int j = t; j < B; j += kThreads)
  {
    float x0, x1;
    if (j < half) { x0 = fwd_sample(job.first, 2 * j); x1 = fwd_sample(job.first, 2 * j + 1); }
    else { x0 = fwd_sample(job.second, 2 * j - B); x1 = fwd_sample(job.second, 2 * j + 1 - B); }
    a[j] = make_float2(x0, x1);
  }
  __syncthreads();

  // all reads of `first` (= prev) are done: it may now be overwritten
  if (job.save_second)
  {
    for (int j = t + half; j < B; j += kThreads)
    {
      const float2 v = a[j];
      reinterpret_cast<float2*>(job.save_second)[j - half] = v;
    }
  }

  if (job.level)
  {
    float m = 0.0f;
    for (int j = t + half; j < B; j += kThreads) { const float2 v = a[j]; m = fmaxf(m, fmaxf(fabsf(v.x), fabsf(v.y))); }
    m = block_max(m);
    if (t == 0) *job.level = m;
  }

  const float2* z = fft_smem<false, TWS>(a, b, B, tw);

  float2* dst = job.dst;
  if (!dst) dst = job.ring + size_t(new_head) * B;

  // split: X[k] = E[k] + w^k O[k], w = exp(-2 pi i / 2B)
  for (int k = t; k <= half; k += kThreads)
  {
    if (k == 0)
    {
      const float2 z0 = z[0];
      dst[0] = make_float2(z0.x + z0.y, z0.x - z0.y);  // (DC, Nyquist)
    }
    else
    {
      const int k2 = B - k;
      const float2 zk = z[k], zk2 = z[k2];
      const float er = 0.5f * (zk.x + zk2.x), ei = 0.5f * (zk.y - zk2.y);
      const float orr = 0.5f * (zk.y + zk2.y), oi = -0.5f * (zk.x - zk2.x);
      const float2 w = TWS ? tw[k] : __ldg(&tw[k]);
      const float tr = orr * w.x - oi * w.y, ti = orr * w.y + oi * w.x;
      dst[k] = make_float2(er + tr, ei + ti);
      if (k2 != k) dst[k2] = make_float2(er - tr, -(ei - ti));
    }
  }

  // Every thread read the old head before the first __syncthreads above, so
  // thread 0 may publish the new one now.  No other CTA of this launch touches
  // this ring; the next kernel in stream order (MAC) sees the new value.
  if (!DEFER_HEAD && !job.dst && t == 0) *job.head = new_head;
}

// Dynamic renderers: CTA n (= source slot n) also records the two ears' current
// filters of this block in the filter ring (see DynTables; declared below).
struct DynRingEntry;
struct DynCtl;
struct FwdRingUpdate
{
  const DynCtl* ctl;            // null: nothing to do
  const DynRingEntry* upd;      // [sources][2]
  DynRingEntry* ring;           // [sources][2][R]
  int R;
};
__device__ __forceinline__ void fwd_ring_update(const FwdRingUpdate& u, int source, int ear);

// in_base != null: job i reads its new block from in_base + i * B instead of the address in
// the job table -- the renderers pass the caller's page-locked input array (mapped host
// memory) here, so that no H2D copy precedes the launch and, with the overlapped variant
// below, the PCIe transfer itself runs under the MAC.
__global__ void __launch_bounds__(kThreads)
fwd_fft_kernel(const FwdJob* __restrict__ jobs, int B, const float2* __restrict__ tw, const FwdRingUpdate ru,
               const float* __restrict__ in_base)
{
  FwdJob job = jobs[blockIdx.x];
  if (in_base) job.second.ptr = in_base + size_t(blockIdx.x) * B;
  if (ru.ctl && threadIdx.x < 2) fwd_ring_update(ru, blockIdx.x, threadIdx.x);
  fwd_fft_body<true>(job, B, tw);
}

// B > 4096: the two transform buffers take 128 KiB of shared memory; the twiddles stay in global memory
__global__ void __launch_bounds__(kThreads)
fwd_fft_large_kernel(const FwdJob* __restrict__ jobs, int B, const float2* __restrict__ tw, const FwdRingUpdate ru,
                     const float* __restrict__ in_base)
{
  FwdJob job = jobs[blockIdx.x];
  if (in_base) job.second.ptr = in_base + size_t(blockIdx.x) * B;
  if (ru.ctl && threadIdx.x < 2) fwd_ring_update(ru, blockIdx.x, threadIdx.x);
  fwd_fft_body<false>(job, B, tw);
}

// The block loop's forward launch when the MAC behind it overlaps it (Engine::fwd_overlap):
// part of the programmatic-dependent-launch chain inverse(t-1) -> forward(t) -> MAC(t).
// It waits for the previous block's inverse kernel (which commits the ring heads and is
// the last reader of the partial accumulators), THEN lets the MAC start: the MAC's CTAs
// move in next to ours (16 KiB of shared memory here, no staged twiddles, leaves room
// for the MAC's 205 KiB ring) and stream every term that does not need this block's
// spectrum while we compute it.  The ring heads stay untouched (DEFER_HEAD).
// (VARIANT only separates the function attributes: 0 runs next to the bulk-copy MAC under the
// maximal shared-memory carve-out, 1 next to the generated-terms MAC under the default one)
// late_wait (Engine::block_overlap): the previous block's inverse kernel has committed the heads BEFORE it let
// us in (HeadCommit::early_trigger == 2) and nothing we touch -- the input block, `prev`, the ring slot
// head + 1 -- is read by it, so the trigger is passed on at once and the MAC of this block starts while that
// inverse kernel still runs.  We wait for it at our END instead: our completion (which the MAC waits for in
// front of its p == 0 terms and of its first partial-spectra store) still implies the previous block's.
template<int VARIANT>
__global__ void __launch_bounds__(kThreads)
fwd_fft_overlap_kernel(const FwdJob* __restrict__ jobs, int B, const float2* __restrict__ tw, const FwdRingUpdate ru,
                       const float* __restrict__ in_base, int late_wait)
{
  if (!late_wait) grid_dependency_wait();
  grid_launch_dependents();
  FwdJob job = jobs[blockIdx.x];
  if (in_base) job.second.ptr = in_base + size_t(blockIdx.x) * B;
  if (ru.ctl && threadIdx.x < 2) fwd_ring_update(ru, blockIdx.x, threadIdx.x);
  fwd_fft_body<false, true>(job, B, tw);
  if (late_wait) grid_dependency_wait();
}

// Filter-set loading: one launch transforms `channels x nvalid` partitions described
// by ONE descriptor (no per-transform job table): CTA (p, c) transforms taps
// [p*B, p*B + B) of channel c into set[(first_channel + c) * partitions + p].
struct FwdBatch
{
  const float* time;       // time-domain taps (memory mode) or the envelope table (synth mode)
  long frame_stride, channel_stride;
  long frames;             // taps per channel
  float2* set;             // filter set base
  int partitions;          // partitions per channel in the set
  int first_channel;       // channel offset in the set
  int synth;               // 1: generate u(seed, id_offset + c, tap) * envelope[tap] * scale
  float scale;
  uint64_t seed, id_offset;
  uint64_t tap_offset;     // synth: the set holds taps [tap_offset, tap_offset + frames) of the generated responses
};

__global__ void __launch_bounds__(kThreads)
fwd_fft_batch_kernel(const FwdBatch batch, int B, const float2* __restrict__ tw)
{
  const int p = blockIdx.x;
  const long c = blockIdx.y + long(blockIdx.z) * gridDim.y;
  FwdJob job;
  const long first_tap = long(p) * B;
  const long left = batch.frames - first_tap;
  job.first.count = int(left < B ? left : B);
  job.first.synth = batch.synth;
  job.first.scale = batch.scale;
  job.first.pad = 0;
  if (batch.synth)
  {
    job.first.ptr = batch.time;   // envelope (indexed by first_index + i) or null
    job.first.stride = 1;
    job.first.key = synth_key(batch.seed, batch.id_offset + uint64_t(c));
    job.first.first_index = batch.tap_offset + uint64_t(first_tap);
  }
  else
  {
    job.first.ptr = batch.time + c * batch.channel_stride + first_tap * batch.frame_stride;
    job.first.stride = batch.frame_stride;
    job.first.key = 0; job.first.first_index = 0;
  }
  job.second.ptr = nullptr; job.second.stride = 1; job.second.count = 0; job.second.synth = 0;
  job.second.key = 0; job.second.first_index = 0; job.second.scale = 0.f; job.second.pad = 0;
  job.dst = batch.set + (size_t(batch.first_channel + c) * batch.partitions + p) * B;
  job.ring = nullptr; job.head = nullptr; job.parts = 0;
  job.save_second = nullptr; job.level = nullptr;
  fwd_fft_body(job, B, tw);
}

}  // namespace ssrk

#include "mac.cuh"      // MAC kernels
#include "inverse.cuh"  // reduction, inverse + crossfade, output gather
#include "level1.cuh"   // fused Level-1 launch

namespace ssrk
{

// ---------------------------------------------------------------- elementwise
// dst = ca * a + cb * b over one partition spectrum (transform_nested lerp,
// convolver.h:773-817 with src/binauralrenderer.h:372-377); a / b may be null
// (zero-flagged partition).  grid.x = partitions.
struct LerpJob { const float4* a; const float4* b; float4* dst; float ca, cb; };

__global__ void __launch_bounds__(kThreads)
lerp_kernel(const LerpJob* __restrict__ jobs, int nvec)
{
  const LerpJob job = jobs[blockIdx.x];
  const float ca = job.ca, cb = job.cb;
  for (int i = threadIdx.x; i < nvec; i += kThreads)
  {
    float4 r = make_float4(0.f, 0.f, 0.f, 0.f);
    if (job.a) { const float4 v = job.a[i]; r.x = ca * v.x; r.y = ca * v.y; r.z = ca * v.z; r.w = ca * v.w; }
    if (job.b)
    {
      const float4 v = job.b[i];
      r.x = fmaf(cb, v.x, r.x); r.y = fmaf(cb, v.y, r.y);
      r.z = fmaf(cb, v.z, r.z); r.w = fmaf(cb, v.w, r.w);
    }
    job.dst[i] = r;
  }
}

}  // namespace ssrk
