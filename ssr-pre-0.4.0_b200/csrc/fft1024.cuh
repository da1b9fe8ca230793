// fft1024.cuh -- specialised real FFT kernels for B = 1024 (2048-point real
// transforms, the block size of every BASELINE configuration).
//
// ONE WARP PER TRANSFORM.  The 1024-point complex FFT is decomposed 32 x 32:
//
//   n = 32 n1 + n2,  k = k1 + 32 k2
//   X[k1 + 32 k2] = sum_{n2} W_32^{n2 k2} ( W_1024^{n2 k1} sum_{n1} x[32 n1 + n2] W_32^{n1 k1} )
//
//   1. lane n2 loads x[32 n1 + n2], n1 = 0..31 (coalesced) and runs a 32-point
//      FFT over n1 entirely in its registers;
//   2. twiddle by W_1024^{n2 k1};
//   3. one transpose through shared memory (33-float2 row pitch, conflict free,
//      __syncwarp only -- no block barrier anywhere);
//   4. lane k1 runs the second 32-point register FFT over n2 and holds
//      X[k1 + 32 k2] in slot k2: stores are coalesced again.
//
// The real split / merge needs X[k] and X[1024 - k] together; 1024 - k lives in
// lane (32 - k1) mod 32, so the conjugate-symmetric butterflies are done with
// warp shuffles (__shfl_sync) instead of another trip through shared memory.
//
// The generic shared-memory Stockham kernels of kernels.cuh remain for every
// other block size and as the A/B baseline (SSR_B200_FFT=generic).
#pragma once

#include "kernels.cuh"

namespace ssrk
{

constexpr int kWarpsPerCta = 4;      // transforms per CTA (4 x 8.25 KiB of transpose tiles)
constexpr int kPitch = 33;           // float2 row pitch of the transpose tile

__host__ __device__ constexpr int brev5(int x)
{
  return ((x & 1) << 4) | ((x & 2) << 2) | (x & 4) | ((x & 8) >> 2) | ((x & 16) >> 4);
}

// cos / sin of 2 pi j / 32, j = 0..15
#define SSR_C32 { 1.0f, 0.98078528040323043f, 0.92387953251128674f, 0.83146961230254524f, \
  0.70710678118654757f, 0.55557023301960218f, 0.38268343236508978f, 0.19509032201612825f, \
  0.0f, -0.19509032201612825f, -0.38268343236508978f, -0.55557023301960218f, \
  -0.70710678118654757f, -0.83146961230254524f, -0.92387953251128674f, -0.98078528040323043f }
#define SSR_S32 { 0.0f, 0.19509032201612825f, 0.38268343236508978f, 0.55557023301960218f, \
  0.70710678118654757f, 0.83146961230254524f, 0.92387953251128674f, 0.98078528040323043f, \
  1.0f, 0.98078528040323043f, 0.92387953251128674f, 0.83146961230254524f, \
  0.70710678118654757f, 0.55557023301960218f, 0.38268343236508978f, 0.19509032201612825f }

// 32-point FFT in registers, radix-2 decimation in frequency: natural order in,
// BIT-REVERSED order out (slot s holds frequency brev5(s)).  All indices are
// compile-time after unrolling, so v[] stays in registers.
template<bool INVERSE>
__device__ __forceinline__ void fft32_regs(float2 (&v)[32])
{
  constexpr float C[16] = SSR_C32;
  constexpr float S[16] = SSR_S32;
#pragma unroll
  for (int half = 16; half >= 1; half >>= 1)
  {
    const int step = 16 / half;  // twiddle index stride: W_{2 half}^j = W_32^{j step}
#pragma unroll
    for (int g = 0; g < 32; g += 2 * half)
    {
#pragma unroll
      for (int j = 0; j < half; ++j)
      {
        const float2 a = v[g + j], b = v[g + j + half];
        v[g + j] = make_float2(a.x + b.x, a.y + b.y);
        const float2 d = make_float2(a.x - b.x, a.y - b.y);
        const int ti = j * step;  // 0..15
        if (ti == 0) v[g + j + half] = d;
        else if (ti == 8)         // W = -i (forward) / +i (inverse)
          v[g + j + half] = INVERSE ? make_float2(-d.y, d.x) : make_float2(d.y, -d.x);
        else
        {
          const float wr = C[ti], wi = INVERSE ? S[ti] : -S[ti];
          v[g + j + half] = make_float2(d.x * wr - d.y * wi, d.x * wi + d.y * wr);
        }
      }
    }
  }
}

// Complex 1024-point FFT of one warp: on entry lane l holds x[32 s + l] in v[s];
// on exit lane l holds X[l + 32 k2] in v[brev5(k2)] (second stage bit-reversed).
// tile: this warp's 32 x kPitch float2 of shared memory.
// tw[i] = exp(-2 pi i * i / 2048).
template<bool INVERSE>
__device__ __forceinline__ void fft1024_warp(float2 (&v)[32], float2* tile,
                                             const float2* __restrict__ tw, int lane)
{
  fft32_regs<INVERSE>(v);          // over n1; slot s = frequency k1 = brev5(s)
  // twiddle W_1024^{n2 k1} and transpose: tile[k1][n2]
#pragma unroll
  for (int s = 0; s < 32; ++s)
  {
    const int k1 = brev5(s);
    float2 x = v[s];
    if (k1 != 0)
    {
      float2 w = __ldg(&tw[2 * ((lane * k1) & 1023)]);
      if (INVERSE) w.y = -w.y;
      x = cmul(x, w);
    }
    tile[k1 * kPitch + lane] = x;
  }
  __syncwarp();
#pragma unroll
  for (int s = 0; s < 32; ++s) v[s] = tile[lane * kPitch + s];   // lane = k1, slot = n2
  __syncwarp();
  fft32_regs<INVERSE>(v);          // over n2; slot s = frequency k2 = brev5(s)
}

// ------------------------------------------------------------------ forward
__device__ __forceinline__ void fwd1024_warp(const FwdJob& job, float2* tile,
                                             const float2* __restrict__ tw, int lane)
{
  constexpr int B = 1024;
  int new_head = 0;
  if (!job.dst)
  {
    new_head = *job.head + 1;
    if (new_head >= job.parts) new_head = 0;
  }
  // z[n] = x[2n] + i x[2n+1], x = [first | second]; lane l, slot s: n = 32 s + l
  float2 v[32];
  const bool fast = !job.first.synth && !job.second.synth && job.first.stride == 1 && job.second.stride == 1
      && job.first.count == B && job.second.count == B && job.first.ptr && job.second.ptr;
  if (fast)
  {
    const float2* f = reinterpret_cast<const float2*>(job.first.ptr);
    const float2* g = reinterpret_cast<const float2*>(job.second.ptr);
#pragma unroll
    for (int s = 0; s < 16; ++s) v[s] = f[32 * s + lane];
#pragma unroll
    for (int s = 16; s < 32; ++s) v[s] = g[32 * (s - 16) + lane];
  }
  else
  {
#pragma unroll
    for (int s = 0; s < 32; ++s)
    {
      const int n = 32 * s + lane;
      v[s] = (s < 16)
        ? make_float2(fwd_sample(job.first, 2 * n), fwd_sample(job.first, 2 * n + 1))
        : make_float2(fwd_sample(job.second, 2 * n - B), fwd_sample(job.second, 2 * n + 1 - B));
    }
  }
  __syncwarp();  // every lane has read `first` (= prev): it may be overwritten now
  if (job.save_second)
  {
    float2* prev = reinterpret_cast<float2*>(job.save_second);
#pragma unroll
    for (int s = 16; s < 32; ++s) prev[32 * (s - 16) + lane] = v[s];
  }
  if (job.level)
  {
    float m = 0.0f;
#pragma unroll
    for (int s = 16; s < 32; ++s) m = fmaxf(m, fmaxf(fabsf(v[s].x), fabsf(v[s].y)));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if (lane == 0) *job.level = m;
  }

  fft1024_warp<false>(v, tile, tw, lane);

  float2* dst = job.dst ? job.dst : job.ring + size_t(new_head) * B;
  // real split: X[k] = E[k] + w^k O[k], E = (Z[k] + conj Z[N-k]) / 2,
  // O = (Z[k] - conj Z[N-k]) / (2i), w = exp(-2 pi i / 2048), N = 1024.
  // k = lane + 32 k2; N - k is in lane (32 - lane) & 31, slot 31 - k2 (lane 0: 32 - k2).
  const int partner = (32 - lane) & 31;
#pragma unroll
  for (int k2 = 0; k2 < 32; ++k2)
  {
    const float2 zk = v[brev5(k2)];
    // what the partner lane holds for ITS frequency slot 31 - k2 ...
    const float2 mine = v[brev5(31 - k2)];
    float2 zp;
    zp.x = __shfl_sync(0xffffffffu, mine.x, partner);
    zp.y = __shfl_sync(0xffffffffu, mine.y, partner);
    // ... except in lane 0, whose partner is itself at slot (32 - k2) & 31
    if (lane == 0) zp = v[brev5((32 - k2) & 31)];
    const int k = lane + 32 * k2;
    float2 out;
    if (k == 0) out = make_float2(zk.x + zk.y, zk.x - zk.y);  // (DC, Nyquist)
    else
    {
      const float er = 0.5f * (zk.x + zp.x), ei = 0.5f * (zk.y - zp.y);
      const float orr = 0.5f * (zk.y + zp.y), oi = -0.5f * (zk.x - zp.x);
      const float2 w = __ldg(&tw[k]);
      out = make_float2(er + (orr * w.x - oi * w.y), ei + (orr * w.y + oi * w.x));
    }
    dst[k] = out;
  }
  if (!job.dst && lane == 0) *job.head = new_head;
}

__global__ void __launch_bounds__(kWarpsPerCta * 32)
fwd_fft1024_kernel(const FwdJob* __restrict__ jobs, int njobs, const float2* __restrict__ tw)
{
  __shared__ float2 tiles[kWarpsPerCta][32 * kPitch];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int j = blockIdx.x * kWarpsPerCta + warp;
  if (j >= njobs) return;
  const FwdJob job = jobs[j];
  fwd1024_warp(job, tiles[warp], tw, lane);
}

__global__ void __launch_bounds__(kWarpsPerCta * 32)
fwd_fft1024_batch_kernel(const FwdBatch batch, int nparts, long ntransforms, const float2* __restrict__ tw)
{
  __shared__ float2 tiles[kWarpsPerCta][32 * kPitch];
  constexpr int B = 1024;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long idx = long(blockIdx.x) * kWarpsPerCta + warp;   // = c * nparts + p
  if (idx >= ntransforms) return;
  const long c = idx / nparts;
  const int p = int(idx - c * nparts);
  FwdJob job;
  const long first_tap = long(p) * B;
  const long left = batch.frames - first_tap;
  job.first.count = int(left < B ? left : B);
  job.first.synth = batch.synth;
  job.first.scale = batch.scale;
  job.first.pad = 0;
  if (batch.synth)
  {
    job.first.ptr = batch.time;
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
  fwd1024_warp(job, tiles[warp], tw, lane);
}

// ------------------------------------------------------------------ inverse
// One warp per output block: for each of <= 3 entries, sum the split partials,
// merge to the half-size complex spectrum (shuffles), inverse FFT, keep the
// second half, scale, crossfade, accumulate; store B floats once.
__global__ void __launch_bounds__(kWarpsPerCta * 32)
inv_fft1024_kernel(const InvJob* __restrict__ jobs, int njobs, const InvEntry* __restrict__ entries,
                   const float2* __restrict__ partials, const float2* __restrict__ reduced,
                   const float2* __restrict__ tw,
                   const float* __restrict__ fade_out, const float* __restrict__ fade_in)
{
  __shared__ float2 tiles[kWarpsPerCta][32 * kPitch];
  constexpr int B = 1024;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int jidx = blockIdx.x * kWarpsPerCta + warp;
  if (jidx >= njobs) return;
  const InvJob job = jobs[jidx];
  float2* tile = tiles[warp];
  const int partner = (32 - lane) & 31;

  float2 res[16];
#pragma unroll
  for (int q = 0; q < 16; ++q) res[q] = make_float2(0.f, 0.f);

  for (int e = 0; e < job.n_entries; ++e)
  {
    const InvEntry en = entries[job.first_entry + e];
    // A[k], k = lane + 32 s (coalesced), summed over the split partials
    float2 v[32];
    if (reduced)
    {
      const float2* r = reduced + size_t(job.first_entry + e) * B;
#pragma unroll
      for (int s = 0; s < 32; ++s) v[s] = __ldcs(r + 32 * s + lane);
    }
    else
    {
#pragma unroll
      for (int s = 0; s < 32; ++s) v[s] = make_float2(0.f, 0.f);
      const float2* base = partials + (size_t(en.part_first) * en.part_stride + en.part_row) * B;
      const size_t pstride = size_t(en.part_stride) * B;
      for (int c = 0; c < en.part_count; ++c)
      {
        const float2* pc = base + size_t(c) * pstride;
#pragma unroll
        for (int s = 0; s < 32; ++s)
        {
          const float2 x = __ldcs(pc + 32 * s + lane);
          v[s].x += x.x; v[s].y += x.y;
        }
      }
    }
    // Z'[k] = (X[k] + conj X[N-k]) + i w^{-k} (X[k] - conj X[N-k]); N - k is in
    // the partner lane at slot 31 - s (lane 0: slot (32 - s) & 31)
    float2 z[32];
#pragma unroll
    for (int s = 0; s < 32; ++s)
    {
      const float2 xk = v[s];
      const float2 mine = v[31 - s];
      float2 xp;
      xp.x = __shfl_sync(0xffffffffu, mine.x, partner);
      xp.y = __shfl_sync(0xffffffffu, mine.y, partner);
      if (lane == 0) xp = v[(32 - s) & 31];
      const int k = lane + 32 * s;
      if (k == 0) z[s] = make_float2(xk.x + xk.y, xk.x - xk.y);   // (DC, Nyquist)
      else
      {
        const float er = xk.x + xp.x, ei = xk.y - xp.y;
        const float dr = xk.x - xp.x, di = xk.y + xp.y;
        float2 w = __ldg(&tw[k]);
        w.y = -w.y;  // w^{-k}
        z[s] = make_float2(er - (dr * w.y + di * w.x), ei + (dr * w.x - di * w.y));
      }
    }
    // inverse complex FFT: "time" index n = 32 s + lane on input, output index
    // j = lane + 32 j2 in slot brev5(j2).  Input here is indexed k = lane + 32 s,
    // i.e. already n = 32 s + lane with the roles of n1 = s, n2 = lane.
    fft1024_warp<true>(z, tile, tw, lane);
    // second half of the 2048 real samples: j >= 512  <=>  j2 >= 16
#pragma unroll
    for (int q = 0; q < 16; ++q)
    {
      const float2 val = z[brev5(16 + q)];
      const int jj = lane + 32 * q;          // j - 512; samples 2 jj, 2 jj + 1 of the output block
      float f0 = en.scale, f1 = en.scale;
      if (en.fade == 1) { const float2 f = __ldg(reinterpret_cast<const float2*>(fade_out) + jj); f0 *= f.x; f1 *= f.y; }
      else if (en.fade == 2) { const float2 f = __ldg(reinterpret_cast<const float2*>(fade_in) + jj); f0 *= f.x; f1 *= f.y; }
      res[q].x = fmaf(val.x, f0, res[q].x);
      res[q].y = fmaf(val.y, f1, res[q].y);
    }
    __syncwarp();
  }

  float m = 0.0f;
#pragma unroll
  for (int q = 0; q < 16; ++q)
  {
    reinterpret_cast<float2*>(job.out)[lane + 32 * q] = res[q];
    m = fmaxf(m, fmaxf(fabsf(res[q].x), fabsf(res[q].y)));
  }
  if (job.level)
  {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if (lane == 0) *job.level = m;
  }
}

}  // namespace ssrk
