// kernels.cuh -- hand-written sm_100a kernels of the partitioned-convolution path.
//
//   fwd_fft_kernel   batched real->packed-complex FFT of [first half | second half]
//                    (input blocks: [prev | cur]; filter partitions: [taps | 0]).
//                    Replaces FFTW R2HC + _sort_coefficients
//                    (apf/apf/convolver.h:144-148, 176-181, 209-268, 324-375).
//   mac_kernel       frequency-domain delay-line complex multiply-accumulate
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
// Complex FFT of size m (power of two) in shared memory, Stockham autosort,
// radix-4 stages plus one radix-2 stage when log2(m) is odd.  `a` holds the
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
    else
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
template<bool TWS = false>
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
    new_head = *job.head + 1;
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
  if (!job.dst && t == 0) *job.head = new_head;
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

__global__ void __launch_bounds__(kThreads)
fwd_fft_kernel(const FwdJob* __restrict__ jobs, int B, const float2* __restrict__ tw, const FwdRingUpdate ru)
{
  const FwdJob job = jobs[blockIdx.x];
  if (ru.ctl && threadIdx.x < 2) fwd_ring_update(ru, blockIdx.x, threadIdx.x);
  fwd_fft_body<true>(job, B, tw);
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
    job.first.first_index = uint64_t(first_tap);
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

// ---------------------------------------------------------------- MAC
// A "term" is one partition-MAC shared by the MT rows of a group:
//   acc[row j] += w * X[input, partition p blocks ago] * H_j
struct MacTermMeta { int input; int p; int wslot; int pad; };

// Dynamic renderers (Binaural / BRS): the term list is not uploaded, it is GENERATED
// by the MAC kernel from a few hundred bytes of per-block control data.
//
// Staggered filter switch (apf/apf/convolver.h:618-699): partition p of a filter
// installed in block b_set becomes active in block b_set + p.  Equivalently, in
// block b partition p uses "the filter that was current in block b - p".  So the
// device keeps, per (source, ear), a ring of the current filter by block index;
// the state after this block's switch reads ring[b - p], the state before it
// ("old", crossfaded out) reads ring[b - 1 - p].
struct DynRingEntry
{
  const float4* base;            // partition 0 of the filter channel (null: empty, convolver.h:415)
  const unsigned char* flags;    // per partition: 1 = carries data (0: fft_node::zero); null = all of them
  int nparts;                    // partitions of that filter; beyond -> empty (convolver.h:650-656)
  int pad;
};

struct DynCtl                    // uploaded once per block
{
  long long b;                   // block index of the state AFTER this block's switch
  int n[3];                      // sources in the constant / old / new accumulator group
  int pad;
};

struct DynTables
{
  const DynCtl* ctl;
  const int* lists;              // [3][nsrc] source slots per group
  const DynRingEntry* ring;      // [nsrc][2 ears][R]
  const int* src_input;          // [nsrc] input id (X ring) of a source slot
  const int* src_parts;          // [nsrc] partitions of its delay line
  unsigned long long* stats;     // [2]: H partitions read, X partitions read (or null)
  int nsrc, R, Pmax;
};

// CTAs [first, first + count) of a G-CTA launch work on group k; proportional to
// the group sizes, at least one CTA per non-empty group.  Every CTA (and the
// reduction that follows) evaluates this identically from the control block.
constexpr int kDynMinTermsPerCta = 4;  // small problems use fewer CTAs (fewer partials to sum)

__host__ __device__ __forceinline__ void dyn_split(const int* n, int Pmax, int G, int k, int& first, int& count)
{
  const int tot = n[0] + n[1] + n[2];
  const int A = (n[0] > 0) + (n[1] > 0) + (n[2] > 0);
  first = 0; count = 0;
  int off = 0;
#pragma unroll
  for (int q = 0; q < 3; ++q)
  {
    int g = n[q] > 0 ? 1 + int((long long)(G - A) * n[q] / tot) : 0;
    const int cap = (n[q] * Pmax + kDynMinTermsPerCta - 1) / kDynMinTermsPerCta;
    if (g > cap) g = cap;
    if (q == k) { first = off; count = g; }
    off += g;
  }
}

struct MacParams
{
  const MacTermMeta* term_meta;      // [nterms]
  const float4* const* term_h;       // [nterms][MT] partition spectra (never null:
                                     //  "empty" partitions point at a zero partition)
  const int4* work;                  // per CTA: x = term_begin, y = term_end, z = partial index
  const float* wtab;                 // weights by slot
  const float4* const* xring;        // per input: ring base
  const int* xparts;                 // per input: ring length
  const int* heads;                  // per input: ring head
  float4* partials;                  // [npartials][MT][B/2] float4
  int B;
  int wtab_offset;                   // added to wslot (selects the weight "kind")
  int partial_offset;                // added to work.z
  int l2_hints;                      // 1: evict-first for H, evict-last for X
  const float4* zero;                // an all-zero partition
  DynTables dyn;                     // mac_kernel<..., DYN = true>: generated terms
};

constexpr int kMacTile = 64;

// STAGES = depth of the register software pipeline (loads of term i + STAGES - 1
// are issued before term i is consumed); MINB = CTAs per SM asked of ptxas.
// DYN: terms generated from MacParams::dyn (dynamic renderers, MT == 2 = the ears)
template<int MT, int NV, int STAGES = 2, int MINB = 1, bool DYN = false>
__global__ void __launch_bounds__(kThreads, MINB)
mac_kernel(const MacParams p)
{
  __shared__ const float4* s_x[kMacTile];
  __shared__ const float4* s_h[kMacTile][MT];
  __shared__ float s_w[kMacTile];
  __shared__ int s_count;

  grid_dependency_wait();
  const int t = threadIdx.x;
  const int lane = t & 31;
  const int nvec = p.B >> 1;  // float4 per partition
  static_assert(!DYN || MT == 2, "generated terms: rows are the two ears");
  int4 wk;
  int kind = 0;
  if constexpr (DYN)
  {
    // generated work: which group this CTA serves and its slice of the group's
    // (source, partition) terms
    const int n0 = p.dyn.ctl->n[0], n1 = p.dyn.ctl->n[1], n2 = p.dyn.ctl->n[2];
    const int nn[3] = {n0, n1, n2};
    int first = 0, count = 0;
    kind = -1;
#pragma unroll
    for (int k = 0; k < 3; ++k)
    {
      int f, c;
      dyn_split(nn, p.dyn.Pmax, gridDim.x, k, f, c);
      if (int(blockIdx.x) >= f && int(blockIdx.x) < f + c) { kind = k; first = f; count = c; }
    }
    if (kind < 0) return;
    const int terms = nn[kind] * p.dyn.Pmax;
    const int chunk = (terms + count - 1) / count;
    const int s = int(blockIdx.x) - first;
    wk = make_int4(min(terms, s * chunk), min(terms, (s + 1) * chunk), int(blockIdx.x), 0);
  }
  else wk = p.work[blockIdx.x];
  // B > 1024: blockIdx.y selects a 1024-bin tile of the spectrum
  const int tile_off = blockIdx.y * (NV * kThreads);
  const int tt = tile_off + t;
  const uint64_t pol_h = p.l2_hints ? l2_policy_evict_first() : l2_policy_evict_normal();
  const uint64_t pol_x = p.l2_hints ? l2_policy_evict_last() : l2_policy_evict_normal();

  float4 acc[MT][NV];
  float nyq[MT];
#pragma unroll
  for (int j = 0; j < MT; ++j)
  {
    nyq[j] = 0.0f;
#pragma unroll
    for (int v = 0; v < NV; ++v) acc[j][v] = make_float4(0.f, 0.f, 0.f, 0.f);
  }

  bool valid[NV];
#pragma unroll
  for (int v = 0; v < NV; ++v) valid[v] = (tt + v * kThreads) < nvec;

  for (int tb = wk.x; tb < wk.y; tb += kMacTile)
  {
    const int tile_n = min(kMacTile, wk.y - tb);
    __syncthreads();  // previous tile fully consumed
    if (t < 32)
    {
      // warp 0 stages the tile and compacts away terms with zero weight
      int count = 0;
      for (int base = 0; base < tile_n; base += 32)
      {
        const int i = base + lane;
        bool ok = false;
        MacTermMeta m; float w = 0.f;
        const float4* hgen[2] = {nullptr, nullptr};
        if (i < tile_n)
        {
          if constexpr (DYN)
          {
            {
              const int src = p.dyn.lists[kind * p.dyn.nsrc + (tb + i) / p.dyn.Pmax];
              m.input = p.dyn.src_input[src];
              m.p = (tb + i) % p.dyn.Pmax;
              w = p.wtab[kind * p.dyn.nsrc + src];
              if (m.p < p.dyn.src_parts[src] && w != 0.0f)
              {
                long long time = p.dyn.ctl->b - (kind == 1 ? 1 : 0) - m.p;
                int slot = int(time % p.dyn.R);
                if (slot < 0) slot += p.dyn.R;
#pragma unroll
                for (int ear = 0; ear < 2; ++ear)
                {
                  const DynRingEntry en = p.dyn.ring[(size_t(src) * 2 + ear) * p.dyn.R + slot];
                  if (en.base && m.p < en.nparts && (!en.flags || en.flags[m.p])) hgen[ear] = en.base + size_t(m.p) * nvec;
                }
                ok = hgen[0] || hgen[1];
              }
            }
          }
          else
          {
            m = p.term_meta[tb + i];
            w = p.wtab[m.wslot + p.wtab_offset];
            ok = (w != 0.0f);
          }
        }
        const unsigned mask = __ballot_sync(0xffffffffu, ok);
        if (ok)
        {
          const int pos = count + __popc(mask & ((1u << lane) - 1u));
          const int parts = p.xparts[m.input];
          int slot = p.heads[m.input] - m.p;
          if (slot < 0) slot += parts;
          s_x[pos] = p.xring[m.input] + size_t(slot) * nvec;
          s_w[pos] = w;
          if constexpr (DYN)
          {
            s_h[pos][0] = hgen[0] ? hgen[0] : p.zero;
            s_h[pos][1] = hgen[1] ? hgen[1] : p.zero;
          }
          else
          {
#pragma unroll
            for (int j = 0; j < MT; ++j) s_h[pos][j] = p.term_h[size_t(tb + i) * MT + j];
          }
        }
        count += __popc(mask);
        if (DYN && p.dyn.stats && blockIdx.y == 0)
        {
          // traffic accounting: H and X partitions this CTA reads
          const int nh = __popc(__ballot_sync(0xffffffffu, hgen[0] != nullptr))
              + __popc(__ballot_sync(0xffffffffu, hgen[1] != nullptr));
          if (lane == 0 && nh)
          {
            atomicAdd(p.dyn.stats, (unsigned long long)nh);
            // X[n, p] counts once: the old-state group re-reads it from L2
            if (kind != 1) atomicAdd(p.dyn.stats + 1, (unsigned long long)__popc(mask));
          }
        }
      }
      if (lane == 0) s_count = count;
    }
    __syncthreads();
    const int n = s_count;

    // software pipeline: loads of term i + STAGES - 1 are in flight while term i
    // is consumed (buffer indices are compile-time after unrolling)
    float4 xb[STAGES][NV];
    float4 hb[STAGES][MT][NV];

#define SSR_MAC_LOAD(buf, i) \
    { \
      const float4* xp = s_x[i]; \
      _Pragma("unroll") for (int v = 0; v < NV; ++v) \
        if (valid[v]) xb[buf][v] = ldg_stream_hint(xp + tt + v * kThreads, pol_x); \
      _Pragma("unroll") for (int j = 0; j < MT; ++j) \
      { \
        const float4* hp = s_h[i][j]; \
        _Pragma("unroll") for (int v = 0; v < NV; ++v) \
          if (valid[v]) hb[buf][j][v] = ldg_stream_hint(hp + tt + v * kThreads, pol_h); \
      } \
    }

#define SSR_MAC_FMA(buf, i) \
    { \
      const float w = s_w[i]; \
      _Pragma("unroll") for (int v = 0; v < NV; ++v) \
      { \
        if (valid[v]) \
        { \
          float4 x = xb[buf][v]; \
          x.x *= w; x.y *= w; x.z *= w; x.w *= w; \
          _Pragma("unroll") for (int j = 0; j < MT; ++j) \
          { \
            const float4 h = hb[buf][j][v]; \
            acc[j][v].x = fmaf(x.x, h.x, acc[j][v].x); \
            acc[j][v].x = fmaf(-x.y, h.y, acc[j][v].x); \
            acc[j][v].y = fmaf(x.x, h.y, acc[j][v].y); \
            acc[j][v].y = fmaf(x.y, h.x, acc[j][v].y); \
            acc[j][v].z = fmaf(x.z, h.z, acc[j][v].z); \
            acc[j][v].z = fmaf(-x.w, h.w, acc[j][v].z); \
            acc[j][v].w = fmaf(x.z, h.w, acc[j][v].w); \
            acc[j][v].w = fmaf(x.w, h.z, acc[j][v].w); \
            if (v == 0 && tt == 0) nyq[j] = fmaf(x.y, h.y, nyq[j]); \
          } \
        } \
      } \
    }

    if constexpr (STAGES == 2)
    {
      // hand-scheduled double buffer (measurably better code than the generic
      // loop below: 6.3 vs 5.1 TB/s at one CTA per SM)
      if (n > 0) SSR_MAC_LOAD(0, 0)
      int i = 0;
      for (; i + 1 < n; i += 2)
      {
        SSR_MAC_LOAD(1, i + 1)
        SSR_MAC_FMA(0, i)
        if (i + 2 < n) SSR_MAC_LOAD(0, i + 2)
        SSR_MAC_FMA(1, i + 1)
      }
      if (i < n) SSR_MAC_FMA(0, i)
    }
    else
    {
#pragma unroll
    for (int st = 0; st < STAGES - 1; ++st)
      if (st < n) SSR_MAC_LOAD(st, st)
    for (int i = 0; i < n; i += STAGES)
    {
#pragma unroll
      for (int st = 0; st < STAGES; ++st)
      {
        const int curi = i + st;
        if (curi < n)
        {
          const int nxt = curi + STAGES - 1;
          if (nxt < n) SSR_MAC_LOAD((st + STAGES - 1) % STAGES, nxt)
          SSR_MAC_FMA(st, curi)
        }
      }
    }
    }
#undef SSR_MAC_LOAD
#undef SSR_MAC_FMA
  }

  // bin 0 holds (DC, Nyquist), both real products (convolver.h:510-541):
  // acc.x = sum(xr*hr - xi*hi) -> DC = acc.x + nyq ; Nyquist = nyq
  if (tt == 0)
  {
#pragma unroll
    for (int j = 0; j < MT; ++j)
    {
      acc[j][0].x += nyq[j];
      acc[j][0].y = nyq[j];
    }
  }

  float4* out = p.partials + size_t(wk.z + p.partial_offset) * MT * nvec;
#pragma unroll
  for (int j = 0; j < MT; ++j)
#pragma unroll
    for (int v = 0; v < NV; ++v)
      if (valid[v]) out[size_t(j) * nvec + tt + v * kThreads] = acc[j][v];
}

// ---------------------------------------------------------------- MAC, bulk-copy staged
// Variant of mac_kernel for the static (table-driven) plans at B >= 1024: a producer
// warp streams the X tile and the MT H tiles of every term into a ring of STAGES
// shared-memory stages with cp.async.bulk (the TMA unit's 1-D bulk copy) completing
// on mbarriers; the 8 consumer warps read them back with LDS.128.  No registers are
// spent on load staging and up to (STAGES - 1) x (1 + MT) x 8 KiB stay in flight per
// SM however the consumers are scheduled.  Selected with ssr_engine_config::
// mac_variant = 6 (4 stages) / 7 (5 stages); see profiles/ for the A/B.
__device__ __forceinline__ unsigned smem_u32(const void* p)
{
  return static_cast<unsigned>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, unsigned count)
{
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar)
{
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, unsigned bytes)
{
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, unsigned parity)
{
  unsigned done;
  do
  {
    asm volatile("{\n\t.reg .pred p;\n\t"
                 "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
                 "selp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  } while (!done);
}
__device__ __forceinline__ void bulk_copy_g2s(void* dst_smem, const void* src_gmem, unsigned bytes, uint64_t* bar,
                                              uint64_t policy)
{
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
               :: "r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)), "l"(policy) : "memory");
}

constexpr int kTmaThreads = kThreads + 32;     // 8 consumer warps + 1 producer warp
constexpr int kTmaTileVec = 2 * kThreads;      // float4 per 1024-bin tile (8 KiB)
constexpr int kTmaTileBytes = kTmaTileVec * 16;

template<int MT, int STAGES>
constexpr size_t mac_tma_smem_bytes()
{
  return size_t(STAGES) * (1 + MT) * kTmaTileBytes + size_t(STAGES) * (8 + 8 + 4 + 4) + 128;
}

template<int MT, int STAGES>
__global__ void __launch_bounds__(kTmaThreads, 1)
mac_tma_kernel(const MacParams p)
{
  extern __shared__ __align__(128) unsigned char tma_smem[];
  __shared__ const float4* q_x[kMacTile];          // producer-private staging of a term tile
  __shared__ const float4* q_h[kMacTile][MT];
  __shared__ float q_w[kMacTile];
  float4* buf = reinterpret_cast<float4*>(tma_smem);
  uint64_t* full = reinterpret_cast<uint64_t*>(tma_smem + size_t(STAGES) * (1 + MT) * kTmaTileBytes);
  uint64_t* empty = full + STAGES;
  volatile float* stage_w = reinterpret_cast<volatile float*>(empty + STAGES);
  volatile int* stage_end = reinterpret_cast<volatile int*>(const_cast<float*>(stage_w) + STAGES);

  const int t = threadIdx.x;
  const int lane = t & 31;
  const int nvec = p.B >> 1;
  const int tile_off = blockIdx.y * kTmaTileVec;   // B > 1024: blockIdx.y selects a 1024-bin tile

  if (t == 0)
  {
#pragma unroll
    for (int s = 0; s < STAGES; ++s) { mbar_init(full + s, 1); mbar_init(empty + s, kThreads / 32); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncthreads();
  grid_dependency_wait();
  const int4 wk = p.work[blockIdx.x];

  if (t >= kThreads)
  {
    // ---- producer warp: resolve the terms tile by tile, then one lane feeds the ring
    const uint64_t pol_h = p.l2_hints ? l2_policy_evict_first() : l2_policy_evict_normal();
    const uint64_t pol_x = p.l2_hints ? l2_policy_evict_last() : l2_policy_evict_normal();
    int s = 0;
    unsigned phase = 0;
    for (int tb = wk.x; tb < wk.y; tb += kMacTile)
    {
      const int tile_n = min(kMacTile, wk.y - tb);
      int count = 0;
      for (int base = 0; base < tile_n; base += 32)
      {
        const int i = base + lane;
        bool ok = false;
        MacTermMeta m; float w = 0.f;
        if (i < tile_n)
        {
          m = p.term_meta[tb + i];
          w = p.wtab[m.wslot + p.wtab_offset];
          ok = (w != 0.0f);   // a muted source costs no traffic
        }
        const unsigned mask = __ballot_sync(0xffffffffu, ok);
        if (ok)
        {
          const int pos = count + __popc(mask & ((1u << lane) - 1u));
          const int parts = p.xparts[m.input];
          int slot = p.heads[m.input] - m.p;
          if (slot < 0) slot += parts;
          q_x[pos] = p.xring[m.input] + size_t(slot) * nvec;
          q_w[pos] = w;
#pragma unroll
          for (int j = 0; j < MT; ++j) q_h[pos][j] = p.term_h[size_t(tb + i) * MT + j];
        }
        count += __popc(mask);
      }
      __syncwarp();
      if (lane == 0)
      {
        for (int i = 0; i < count; ++i)
        {
          mbar_wait(empty + s, phase ^ 1u);          // the consumers have released this stage
          stage_w[s] = q_w[i];
          stage_end[s] = 0;
          mbar_arrive_expect_tx(full + s, (1 + MT) * kTmaTileBytes);
          float4* dst = buf + size_t(s) * (1 + MT) * kTmaTileVec;
          bulk_copy_g2s(dst, q_x[i] + tile_off, kTmaTileBytes, full + s, pol_x);
#pragma unroll
          for (int j = 0; j < MT; ++j)
            bulk_copy_g2s(dst + size_t(1 + j) * kTmaTileVec, q_h[i][j] + tile_off, kTmaTileBytes, full + s, pol_h);
          if (++s == STAGES) { s = 0; phase ^= 1u; }
        }
      }
      __syncwarp();
    }
    if (lane == 0)
    {
      mbar_wait(empty + s, phase ^ 1u);
      stage_end[s] = 1;                               // no more terms
      mbar_arrive(full + s);
    }
    return;
  }

  // ---- consumer warps
  float4 acc[MT][2];
  float nyq[MT];
#pragma unroll
  for (int j = 0; j < MT; ++j)
  {
    nyq[j] = 0.0f;
    acc[j][0] = make_float4(0.f, 0.f, 0.f, 0.f);
    acc[j][1] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  const int tt = tile_off + t;
  int s = 0;
  unsigned phase = 0;
  while (true)
  {
    mbar_wait(full + s, phase);
    if (stage_end[s]) break;
    const float w = stage_w[s];
    const float4* sb = buf + size_t(s) * (1 + MT) * kTmaTileVec;
#pragma unroll
    for (int v = 0; v < 2; ++v)
    {
      float4 x = sb[t + v * kThreads];
      x.x *= w; x.y *= w; x.z *= w; x.w *= w;
#pragma unroll
      for (int j = 0; j < MT; ++j)
      {
        const float4 h = sb[size_t(1 + j) * kTmaTileVec + t + v * kThreads];
        acc[j][v].x = fmaf(x.x, h.x, acc[j][v].x);
        acc[j][v].x = fmaf(-x.y, h.y, acc[j][v].x);
        acc[j][v].y = fmaf(x.x, h.y, acc[j][v].y);
        acc[j][v].y = fmaf(x.y, h.x, acc[j][v].y);
        acc[j][v].z = fmaf(x.z, h.z, acc[j][v].z);
        acc[j][v].z = fmaf(-x.w, h.w, acc[j][v].z);
        acc[j][v].w = fmaf(x.z, h.w, acc[j][v].w);
        acc[j][v].w = fmaf(x.w, h.z, acc[j][v].w);
        if (v == 0 && tt == 0) nyq[j] = fmaf(x.y, h.y, nyq[j]);
      }
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(empty + s);            // this warp is done with the stage
    if (++s == STAGES) { s = 0; phase ^= 1u; }
  }

  // bin 0 holds (DC, Nyquist), both real products (convolver.h:510-541)
  if (tt == 0)
  {
#pragma unroll
    for (int j = 0; j < MT; ++j) { acc[j][0].x += nyq[j]; acc[j][0].y = nyq[j]; }
  }
  float4* out = p.partials + size_t(wk.z + p.partial_offset) * MT * nvec;
#pragma unroll
  for (int j = 0; j < MT; ++j)
#pragma unroll
    for (int v = 0; v < 2; ++v) out[size_t(j) * nvec + tt + v * kThreads] = acc[j][v];
}

// ---------------------------------------------------------------- inverse
struct InvEntry
{
  int part_first;   // first partial index
  int part_count;   // number of partials to sum (split count of the group)
  int part_stride;  // MT of the group
  int part_row;     // row within the group
  int fade;         // 0 none, 1 fade-out, 2 fade-in (combine_channels.h:297-335)
  float scale;      // 1 / (2B) (convolver.h:458) [times a per-row weight]
};

struct InvJob
{
  int first_entry;
  int n_entries;    // 0 -> output block is zeros (combine_channels.h:126-129)
  float* out;       // B floats (device)
  float* level;     // if non-null: max |out| (src/rendererbase.h:468-472)
};

constexpr int kInvMaxRes = 8;  // B <= 4096: (B/2 float2) / 256 threads

// Multi-GPU output gather fused into the inverse kernel: the output
// blocks of a rank are stored straight into the host-facing rank's gather buffer
// over NVLink peer memory (InvJob::out + out_offset points into it), and every
// CTA then bumps a system-scope arrival counter in that buffer's control block.
// Two slots alternate by block parity; before storing block t a CTA checks that
// the root has released block t-2 (`consumed >= need_consumed`).  All pointers
// null / zero: plain local stores.
struct GatherArgs
{
  long long out_offset;                    // floats added to every InvJob::out (slot * slot stride)
  const unsigned long long* consumed;      // root's "blocks released" counter (peer memory), or null
  unsigned long long need_consumed;
  unsigned long long* arrived;             // this rank's arrival counter on the root (peer memory), or null
  int* error;                              // local sticky flag: a wait timed out
};

constexpr unsigned long long kGatherTimeoutNs = 20ull * 1000ull * 1000ull * 1000ull;

__device__ __forceinline__ unsigned long long globaltimer_ns()
{
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}

__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p)
{
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}

// spins until *p >= target; gives up (sticky error flag) after kGatherTimeoutNs so
// that a dead peer can never hang the GPU
__device__ __noinline__ void gather_wait_ge(const unsigned long long* p, unsigned long long target, int* error)
{
  if (ld_acquire_sys(p) >= target) return;
  if (error && *reinterpret_cast<volatile int*>(error)) return;
  const unsigned long long t0 = globaltimer_ns();
  while (ld_acquire_sys(p) < target)
  {
    __nanosleep(200);
    if (globaltimer_ns() - t0 > kGatherTimeoutNs)
    {
      if (error) *reinterpret_cast<volatile int*>(error) = 1;
      return;
    }
  }
}

// root: wait until every peer rank has delivered its output blocks of block
// `epoch` (0-based): arrived[g] >= (epoch + 1) * ctas[g].  One thread per rank.
struct GatherWaitArgs
{
  const unsigned long long* arrived;   // [world], 16 counters apart (128 bytes)
  const int* ctas;                     // [world] CTAs (= output blocks) per rank; 0 for the root itself
  int world;
  unsigned long long epoch;
  int* error;
};
constexpr int kGatherCounterStride = 16;  // unsigned long longs between counters (one 128-byte line each)

__global__ void gather_wait_kernel(GatherWaitArgs a)
{
  const int g = threadIdx.x;
  if (g < a.world && a.ctas[g] > 0)
    gather_wait_ge(a.arrived + size_t(g) * kGatherCounterStride,
        (a.epoch + 1) * (unsigned long long)a.ctas[g], a.error);
}

// root: block `epoch` has been read out of its slot -> peers may overwrite it
__global__ void gather_release_kernel(unsigned long long* consumed, unsigned long long value)
{
  asm volatile("st.release.sys.global.u64 [%0], %1;" :: "l"(consumed), "l"(value) : "memory");
}

// Parallel pre-reduction of the split partial accumulators, used when there are
// few output blocks but deep splits (Binaural / BRS: 2 outputs; sharded Generic):
// grid (entries, (B/2)/32); warp w of a CTA sums partials w, w+8, ... for a
// 64-bin tile, then the 8 warps are combined through shared memory.
__global__ void __launch_bounds__(kThreads)
reduce_partials_kernel(const InvEntry* __restrict__ entries, const float4* __restrict__ partials,
                       float4* __restrict__ reduced, int nvec)
{
  __shared__ float4 s_part[kThreads / 32][32];
  grid_dependency_wait();
  const InvEntry en = entries[blockIdx.x];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int idx = blockIdx.y * 32 + lane;
  float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
  if (idx < nvec)
  {
    const float4* base = partials + (size_t(en.part_first) * en.part_stride + en.part_row) * nvec + idx;
    const size_t pstride = size_t(en.part_stride) * nvec;
#pragma unroll 4
    for (int c = warp; c < en.part_count; c += kThreads / 32)
    {
      const float4 v = __ldcs(base + size_t(c) * pstride);
      s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
    }
  }
  s_part[warp][lane] = s;
  __syncthreads();
  if (warp == 0 && idx < nvec)
  {
#pragma unroll
    for (int w = 1; w < kThreads / 32; ++w)
    {
      const float4 v = s_part[w][lane];
      s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
    }
    reduced[size_t(blockIdx.x) * nvec + idx] = s;
  }
}

// Dynamic renderers: the same pre-reduction with the split derived from the
// control block (see dyn_split); grid (2 ears x 3 kinds, (B/2)/32).  Entry
// ear * 3 + kind of `reduced` receives the sum; inactive kinds are skipped (the
// inverse kernel skips them too).
__global__ void __launch_bounds__(kThreads)
dyn_reduce_kernel(const DynCtl* __restrict__ ctl, int Pmax, int G, const float4* __restrict__ partials,
                  float4* __restrict__ reduced, int nvec)
{
  __shared__ float4 s_part[kThreads / 32][32];
  grid_dependency_wait();
  const int ear = blockIdx.x / 3, kind = blockIdx.x % 3;
  const int nn[3] = {ctl->n[0], ctl->n[1], ctl->n[2]};
  if (nn[kind] == 0) return;
  int first, count;
  dyn_split(nn, Pmax, G, kind, first, count);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int idx = blockIdx.y * 32 + lane;
  float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
  if (idx < nvec)
  {
    const float4* base = partials + (size_t(first) * 2 + ear) * nvec + idx;
    const size_t pstride = size_t(2) * nvec;
#pragma unroll 4
    for (int c = warp; c < count; c += kThreads / 32)
    {
      const float4 v = __ldcs(base + size_t(c) * pstride);
      s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
    }
  }
  s_part[warp][lane] = s;
  __syncthreads();
  if (warp == 0 && idx < nvec)
  {
#pragma unroll
    for (int w = 1; w < kThreads / 32; ++w)
    {
      const float4 v = s_part[w][lane];
      s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
    }
    reduced[size_t(blockIdx.x) * nvec + idx] = s;
  }
}

__device__ __forceinline__ void fwd_ring_update(const FwdRingUpdate& u, int source, int ear)
{
  int slot = int(u.ctl->b % u.R);
  if (slot < 0) slot += u.R;
  const int row = source * 2 + ear;
  u.ring[size_t(row) * u.R + slot] = u.upd[row];
}

// ring[(source, ear)][b mod R] = the filter that is current in block b
__global__ void dyn_ring_update_kernel(const DynCtl* __restrict__ ctl, const DynRingEntry* __restrict__ upd,
                                       DynRingEntry* __restrict__ ring, int rows, int R)
{
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows) return;
  int slot = int(ctl->b % R);
  if (slot < 0) slot += R;
  ring[size_t(i) * R + slot] = upd[i];
}

// One CTA per output block, one kThreads-thread GROUP per accumulator entry of the
// block (constant / fade-out / fade-in, <= kInvMaxGroups): the up to three inverse
// transforms of a crossfading output run side by side instead of one after the
// other; group 0 adds the faded results and stores the block.
constexpr int kInvMaxGroups = 3;

template<int GROUPS>
__global__ void __launch_bounds__(kThreads * GROUPS, GROUPS == 1 ? 4 : 1)
inv_fft_kernel(const InvJob* __restrict__ jobs, const InvEntry* __restrict__ entries,
               const float2* __restrict__ partials, const float2* __restrict__ reduced, int B,
               const float2* __restrict__ tw,
               const float* __restrict__ fade_out, const float* __restrict__ fade_in,
               const GatherArgs ga, const int* __restrict__ kind_counts)
{
  extern __shared__ float2 smem[];
  __shared__ float s_max[kThreads / 32];
  const int g = GROUPS == 1 ? 0 : threadIdx.x / kThreads;  // entry group
  const int t = threadIdx.x - g * kThreads;                // thread within the group
  constexpr int ngroups = GROUPS;
  float2* a = smem + size_t(g) * 2 * B;
  float2* b = a + B;
  const InvJob job = jobs[blockIdx.x];
  const int half = B >> 1;
  // twiddle table -> shared memory, once for all groups
  float2* stw = smem + size_t(GROUPS) * 2 * B;
  stage_twiddles(stw, tw, B, threadIdx.x, kThreads * GROUPS);  // constants: may overlap the predecessor's tail
  grid_dependency_wait();
  __syncthreads();
  // flow control of the fused gather: the slot this block goes to must have been
  // read out by the root (checked by one thread while the others start loading)
  if (ga.consumed && threadIdx.x == 0) gather_wait_ge(ga.consumed, ga.need_consumed, ga.error);

  float2 res[kInvMaxRes];
#pragma unroll
  for (int q = 0; q < kInvMaxRes; ++q) res[q] = make_float2(0.f, 0.f);

  // entries beyond the launched groups are handled by the last group in turn
  for (int e = g; e < job.n_entries; e += ngroups)
  {
    const InvEntry en = entries[job.first_entry + e];
    // dynamic renderers: accumulator kinds without sources this block are skipped
    if (kind_counts && kind_counts[en.fade] == 0) continue;
    // A[k] = sum over split partials
    // (float4 = two bins per thread; the loop over partials is unrolled so that
    // 8 independent loads are in flight per thread)
    if (reduced)
    {
      const float4* r = reinterpret_cast<const float4*>(reduced) + size_t(job.first_entry + e) * (B >> 1);
      for (int k = t; k < (B >> 1); k += kThreads) reinterpret_cast<float4*>(a)[k] = __ldcs(r + k);
    }
    else
    {
      const float4* pbase = reinterpret_cast<const float4*>(partials)
          + (size_t(en.part_first) * en.part_stride + en.part_row) * (B >> 1);
      const size_t pstride = size_t(en.part_stride) * (B >> 1);
      for (int k = t; k < (B >> 1); k += kThreads)
      {
        float4 s0 = make_float4(0.f, 0.f, 0.f, 0.f), s1 = s0;
        int c = 0;
        for (; c + 8 <= en.part_count; c += 8)
        {
          float4 v[8];
#pragma unroll
          for (int u = 0; u < 8; ++u) v[u] = __ldcs(pbase + size_t(c + u) * pstride + k);
#pragma unroll
          for (int u = 0; u < 8; u += 2)
          {
            s0.x += v[u].x; s0.y += v[u].y; s0.z += v[u].z; s0.w += v[u].w;
            s1.x += v[u + 1].x; s1.y += v[u + 1].y; s1.z += v[u + 1].z; s1.w += v[u + 1].w;
          }
        }
        for (; c < en.part_count; ++c)
        {
          const float4 v = __ldcs(pbase + size_t(c) * pstride + k);
          s0.x += v.x; s0.y += v.y; s0.z += v.z; s0.w += v.w;
        }
        reinterpret_cast<float4*>(a)[k] = make_float4(s0.x + s1.x, s0.y + s1.y, s0.z + s1.z, s0.w + s1.w);
      }
    }
    group_sync(g + 1);
    // Z'[k] = (X[k] + conj X[B-k]) + i w^{-k} (X[k] - conj X[B-k])
    for (int k = t; k <= half; k += kThreads)
    {
      if (k == 0)
      {
        const float2 x0 = a[0];  // (DC, Nyquist)
        b[0] = make_float2(x0.x + x0.y, x0.x - x0.y);
      }
      else
      {
        const int k2 = B - k;
        const float2 xk = a[k], xk2 = a[k2];
        const float er = xk.x + xk2.x, ei = xk.y - xk2.y;
        const float dr = xk.x - xk2.x, di = xk.y + xk2.y;
        float2 w = stw[k];
        w.y = -w.y;  // w^{-k}
        const float tr = -(dr * w.y + di * w.x), ti = dr * w.x - di * w.y;
        b[k] = make_float2(er + tr, ei + ti);
        if (k2 != k) b[k2] = make_float2(er - tr, -(ei - ti));
      }
    }
    group_sync(g + 1);
    const float2* z = fft_smem<true, true>(b, a, B, stw, t, g + 1);
    // keep the second half: time index i = 2j, 2j+1 with j >= B/2
#pragma unroll
    for (int q = 0; q < kInvMaxRes; ++q)
    {
      const int jj = t + q * kThreads;  // j - half
      if (jj < half)
      {
        const float2 v = z[half + jj];
        float f0 = en.scale, f1 = en.scale;
        if (en.fade == 1) { f0 *= __ldg(fade_out + 2 * jj); f1 *= __ldg(fade_out + 2 * jj + 1); }
        else if (en.fade == 2) { f0 *= __ldg(fade_in + 2 * jj); f1 *= __ldg(fade_in + 2 * jj + 1); }
        res[q].x = fmaf(v.x, f0, res[q].x);
        res[q].y = fmaf(v.y, f1, res[q].y);
      }
    }
    group_sync(g + 1);  // z (smem) is rewritten by the group's next entry / the hand-over below
  }

  // groups 1.. hand their faded blocks to group 0 through their own smem region
  if (ngroups > 1)
  {
    if (g > 0)
    {
#pragma unroll
      for (int q = 0; q < kInvMaxRes; ++q)
      {
        const int jj = t + q * kThreads;
        if (jj < half) a[jj] = res[q];
      }
    }
    __syncthreads();
    if (g == 0)
      for (int gg = 1; gg < ngroups; ++gg)
      {
        const float2* other = smem + size_t(gg) * 2 * B;
#pragma unroll
        for (int q = 0; q < kInvMaxRes; ++q)
        {
          const int jj = t + q * kThreads;
          if (jj < half) { const float2 v = other[jj]; res[q].x += v.x; res[q].y += v.y; }
        }
      }
  }

  if (ga.consumed) __syncthreads();  // thread 0's slot check precedes every store
  if (g == 0)
  {
    float m = 0.0f;
    float2* out = reinterpret_cast<float2*>(job.out + ga.out_offset);
#pragma unroll
    for (int q = 0; q < kInvMaxRes; ++q)
    {
      const int jj = t + q * kThreads;
      if (jj < half)
      {
        out[jj] = res[q];
        m = fmaxf(m, fmaxf(fabsf(res[q].x), fabsf(res[q].y)));
      }
    }
    if (job.level)
    {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
      if ((t & 31) == 0) s_max[t >> 5] = m;
      group_sync(1);
      if (t < 32)
      {
        m = t < kThreads / 32 ? s_max[t] : 0.0f;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
        if (t == 0) *job.level = m;
      }
    }
  }
  if (ga.arrived)
  {
    // all stores of this CTA (peer memory, over NVLink) -> fence -> arrival count
    __syncthreads();
    if (threadIdx.x == 0)
    {
      __threadfence_system();
      atomicAdd_system(ga.arrived, 1ull);
    }
  }
}

// ---------------------------------------------------------------- Level 1, fused
// One (Input, Output) pair of the apf::conv class API in ONE launch of ONE CTA:
// Input::add_block's forward transform (convolver.h:324-375), OutputBase::convolve's
// spectral MAC over the partitions, inverse transform and weight / n scaling
// (convolver.h:435-465).  For a lone convolver with a few partitions (BASELINE
// config 1: StaticConvolver, 8 partitions) the work is ~10 us, so the cost is
// launches, copies and the final synchronisation: the input block is read straight
// from the caller-side pinned staging buffer (mapped host memory), the result block
// and a completion flag are written straight back to mapped host memory, and the
// host spins on the flag instead of synchronising the stream.
struct FusedParams
{
  FwdJob fwd;                        // forward job of the Input (second half = mapped host block)
  int do_fwd;                        // 0: the block was already transformed (another Output came first)
  const MacTermMeta* term_meta;      // the Output's term list (one row)
  const float4* const* term_h;
  int nterms;
  const float4* const* xring;
  const int* xparts;
  const int* heads;
  float scale;                       // weight / (2B)
  float* out;                        // B floats, mapped host memory
  int* flag;                         // mapped host memory; <- seq when `out` is complete
  int seq;
  int B;
};

constexpr int kFusedMaxTerms = 64;

template<int NV>
__global__ void __launch_bounds__(kThreads)
level1_fused_kernel(const FusedParams p, const float2* __restrict__ tw_global)
{
  extern __shared__ float2 smem[];
  __shared__ const float4* s_x[kFusedMaxTerms];
  __shared__ const float4* s_h[kFusedMaxTerms];
  const int B = p.B;
  const int t = threadIdx.x;
  const int half = B >> 1;
  const int nvec = B >> 1;
  float2* a = smem;
  float2* b = smem + B;
  const float2* stw = smem + 2 * B;

  if (p.do_fwd) fwd_fft_body<true>(p.fwd, B, tw_global);   // stages the twiddles, advances the ring head
  else stage_twiddles(smem + 2 * B, tw_global, B, t, kThreads);
  __syncthreads();  // new spectrum + head (global, this CTA's own writes) and twiddles visible

  // resolve the term pointers (plain loads: the head and X[p = 0] were written above)
  if (t < p.nterms)
  {
    const MacTermMeta m = p.term_meta[t];
    const int parts = p.xparts[m.input];
    int slot = *reinterpret_cast<const volatile int*>(p.heads + m.input) - m.p;
    if (slot < 0) slot += parts;
    s_x[t] = p.xring[m.input] + size_t(slot) * nvec;
    s_h[t] = p.term_h[t];
  }
  __syncthreads();

  float4 acc[NV];
#pragma unroll
  for (int v = 0; v < NV; ++v) acc[v] = make_float4(0.f, 0.f, 0.f, 0.f);
  float nyq = 0.0f;
  bool valid[NV];
#pragma unroll
  for (int v = 0; v < NV; ++v) valid[v] = (t + v * kThreads) < nvec;

  // U terms per step so that ~16 independent 16-byte loads are in flight per thread
  constexpr int U = NV >= 8 ? 1 : 8 / NV;
  for (int i0 = 0; i0 < p.nterms; i0 += U)
  {
    float4 xb[U][NV], hb[U][NV];
#pragma unroll
    for (int u = 0; u < U; ++u)
      if (i0 + u < p.nterms)
      {
        const float4* xp = s_x[i0 + u];
        const float4* hp = s_h[i0 + u];
#pragma unroll
        for (int v = 0; v < NV; ++v)
          if (valid[v])
          {
            xb[u][v] = xp[t + v * kThreads];  // coherent load (X[p = 0] was written by this CTA)
            hb[u][v] = ldg_stream(hp + t + v * kThreads);
          }
      }
#pragma unroll
    for (int u = 0; u < U; ++u)
      if (i0 + u < p.nterms)
      {
#pragma unroll
        for (int v = 0; v < NV; ++v)
          if (valid[v])
          {
            const float4 x = xb[u][v], h = hb[u][v];
            acc[v].x = fmaf(x.x, h.x, acc[v].x);
            acc[v].x = fmaf(-x.y, h.y, acc[v].x);
            acc[v].y = fmaf(x.x, h.y, acc[v].y);
            acc[v].y = fmaf(x.y, h.x, acc[v].y);
            acc[v].z = fmaf(x.z, h.z, acc[v].z);
            acc[v].z = fmaf(-x.w, h.w, acc[v].z);
            acc[v].w = fmaf(x.z, h.w, acc[v].w);
            acc[v].w = fmaf(x.w, h.z, acc[v].w);
            if (v == 0 && t == 0) nyq = fmaf(x.y, h.y, nyq);
          }
      }
  }
  // bin 0 holds (DC, Nyquist), both real products (convolver.h:510-541)
  if (t == 0) { acc[0].x += nyq; acc[0].y = nyq; }
#pragma unroll
  for (int v = 0; v < NV; ++v)
    if (valid[v]) reinterpret_cast<float4*>(a)[t + v * kThreads] = acc[v];
  __syncthreads();

  // inverse real split, complex inverse FFT, second half * weight / n
  for (int k = t; k <= half; k += kThreads)
  {
    if (k == 0)
    {
      const float2 x0 = a[0];
      b[0] = make_float2(x0.x + x0.y, x0.x - x0.y);
    }
    else
    {
      const int k2 = B - k;
      const float2 xk = a[k], xk2 = a[k2];
      const float er = xk.x + xk2.x, ei = xk.y - xk2.y;
      const float dr = xk.x - xk2.x, di = xk.y + xk2.y;
      float2 w = stw[k];
      w.y = -w.y;
      const float tr = -(dr * w.y + di * w.x), ti = dr * w.x - di * w.y;
      b[k] = make_float2(er + tr, ei + ti);
      if (k2 != k) b[k2] = make_float2(er - tr, -(ei - ti));
    }
  }
  __syncthreads();
  const float2* z = fft_smem<true, true>(b, a, B, stw);
  for (int jj = t; jj < half; jj += kThreads)
  {
    const float2 v = z[half + jj];
    reinterpret_cast<float2*>(p.out)[jj] = make_float2(v.x * p.scale, v.y * p.scale);
  }
  __syncthreads();
  if (t == 0)
  {
    __threadfence_system();
    *reinterpret_cast<volatile int*>(p.flag) = p.seq;
  }
}

// ---------------------------------------------------------------- elementwise
// dst = ca * a + cb * b over one partition spectrum (transform_nested lerp,
// convolver.h:773-817 with src/binauralrenderer.h:372-377); a / b may be null
// (zero-flagged partition).  grid.x = partitions.
struct LerpJob { const float4* a; const float4* b; float4* dst; };

__global__ void __launch_bounds__(kThreads)
lerp_kernel(const LerpJob* __restrict__ jobs, int nvec, float ca, float cb)
{
  const LerpJob job = jobs[blockIdx.x];
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
