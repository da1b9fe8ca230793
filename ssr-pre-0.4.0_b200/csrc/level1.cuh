// level1.cuh -- one (Input, Output) pair of the apf::conv class API in one launch
// (part of the kernel set of kernels.cuh; included from there, in this order)
#pragma once

namespace ssrk
{

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

}  // namespace ssrk
