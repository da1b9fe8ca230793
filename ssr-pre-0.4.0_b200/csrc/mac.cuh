// mac.cuh -- the spectral multiply-accumulate kernels: register-staged (table-driven or generated terms) and bulk-copy (TMA) staged
// (part of the kernel set of kernels.cuh; included from there, in this order)
#pragma once

namespace ssrk
{

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
  const DynRingEntry* upd;       // [nsrc][2] the filters current in THIS block (= what the forward
                                 //  kernel stores into ring[b]); read instead of ring[b], so that
                                 //  nothing here depends on the forward kernel's ring update
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
  const int4* work;                  // SEGMENTS: x = term_begin, y = term_end, z = partial index; a segment
                                     //  never crosses a group boundary
  const int* cta_first;              // [CTAs + 1]: CTA c works through segments [cta_first[c], cta_first[c + 1])
                                     //  (usually one; two or more where its share of the term list straddles
                                     //  a group boundary -- that is what lets 16 groups fill all 148 SMs)
  const int4* cta_head;              // bulk-copy MAC: per CTA {first segment, end segment, 0, 0} and a copy of its
                                     //  first segment (work[first]) -- ONE load level from blockIdx to the first
                                     //  term tile instead of two (the start-up chain of a 25 - 50 us launch)
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
  unsigned long long* trace;         // per CTA {start, first data, end, accumulated busy time} in %globaltimer ns, or
                                     //  null (bulk-copy MAC): feeds the share calibration (Renderer::calibrate)
  int early_trigger;                 // 1: griddepcontrol.launch_dependents at the start of the bulk-copy MACs
  int trace_mode;                    // 0: trace[3] accumulates the CTA's busy time (share calibration);
                                     //  1: trace[3] = time of the producer's first bulk copy (scripts/mac_trace.py)
  int head_bias;                     // 1: the forward transform of this block runs CONCURRENTLY (it has not
                                     //  published the ring heads yet): newest slot = heads[] + 1, and terms with
                                     //  p == 0 are issued last, behind griddepcontrol.wait (mac_tma_kernel)
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
  // generated terms with the forward transform still running (head_bias): terms with
  // p == 0 -- the only ones that read the spectrum it is producing -- are held back here
  __shared__ const float4* s_lx[DYN ? kMacTile : 1];
  __shared__ const float4* s_lh[DYN ? kMacTile : 1][MT];
  __shared__ float s_lw[DYN ? kMacTile : 1];
  __shared__ int s_late;

  const bool defer = DYN && p.head_bias != 0;
  bool waited = !defer;
  if (!defer) grid_dependency_wait();
  if (DYN && threadIdx.x == 0) s_late = 0;
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
  int seg = 0, seg_end = 1;
  if constexpr (!DYN) { seg = p.cta_first[blockIdx.x]; seg_end = p.cta_first[blockIdx.x + 1]; }
  // B > 1024: blockIdx.y selects a 1024-bin tile of the spectrum
  const int tile_off = blockIdx.y * (NV * kThreads);
  const int tt = tile_off + t;
  const uint64_t pol_h = p.l2_hints ? l2_policy_evict_first() : l2_policy_evict_normal();
  const uint64_t pol_x = p.l2_hints ? l2_policy_evict_last() : l2_policy_evict_normal();

  bool valid[NV];
#pragma unroll
  for (int v = 0; v < NV; ++v) valid[v] = (tt + v * kThreads) < nvec;

  for (; seg < seg_end; ++seg)
  {
  if constexpr (!DYN) wk = p.work[seg];
  float4 acc[MT][NV];
  float nyq[MT];
#pragma unroll
  for (int j = 0; j < MT; ++j)
  {
    nyq[j] = 0.0f;
#pragma unroll
    for (int v = 0; v < NV; ++v) acc[j][v] = make_float4(0.f, 0.f, 0.f, 0.f);
  }

  for (int tb = wk.x; ; tb += kMacTile)
  {
    const bool last_pass = tb >= wk.y;   // generated terms: one more pass for the held-back terms
    if (last_pass && !DYN) break;
    const int tile_n = last_pass ? 0 : min(kMacTile, wk.y - tb);
    __syncthreads();  // previous tile fully consumed
    if constexpr (DYN)
    {
      // (uniform over the CTA: s_late is only written between barriers)
      if (!waited && (last_pass || s_late + tile_n > kMacTile))
      {
        grid_dependency_wait();          // the forward transform of this block is complete
        waited = true;
      }
      if (last_pass)
      {
        const int n_late = s_late;
        if (n_late == 0) break;
        if (t < n_late)
        {
          s_x[t] = s_lx[t]; s_w[t] = s_lw[t];
#pragma unroll
          for (int j = 0; j < MT; ++j) s_h[t][j] = s_lh[t][j];
        }
        if (t == 0) s_count = n_late;
      }
    }
    if (t < 32 && !last_pass)
    {
      // warp 0 stages the tile and compacts away terms with zero weight
      int count = 0;
      int late_count = DYN ? s_late : 0;
      for (int base = 0; base < tile_n; base += 32)
      {
        const int i = base + lane;
        bool ok = false, late = false;
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
                const bool current = (kind != 1 && m.p == 0);   // ring[b] = this block's update
#pragma unroll
                for (int ear = 0; ear < 2; ++ear)
                {
                  const DynRingEntry en = current ? p.dyn.upd[size_t(src) * 2 + ear]
                                                  : p.dyn.ring[(size_t(src) * 2 + ear) * p.dyn.R + slot];
                  if (en.base && m.p < en.nparts && (!en.flags || en.flags[m.p])) hgen[ear] = en.base + size_t(m.p) * nvec;
                }
                ok = hgen[0] || hgen[1];
                late = ok && !waited && m.p == 0;
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
        const unsigned mask = __ballot_sync(0xffffffffu, ok && !late);
        const unsigned mask_late = DYN ? __ballot_sync(0xffffffffu, late) : 0u;
        if (ok)
        {
          const int parts = p.xparts[m.input];
          int slot = p.heads[m.input] + p.head_bias - m.p;
          if (slot < 0) slot += parts;
          if (slot >= parts) slot -= parts;
          const float4* xptr = p.xring[m.input] + size_t(slot) * nvec;
          if (DYN && late)
          {
            const int lpos = late_count + __popc(mask_late & ((1u << lane) - 1u));
            s_lx[lpos] = xptr;
            s_lw[lpos] = w;
            s_lh[lpos][0] = hgen[0] ? hgen[0] : p.zero;
            s_lh[lpos][MT - 1] = hgen[1] ? hgen[1] : p.zero;
          }
          const int pos = count + __popc(mask & ((1u << lane) - 1u));
          if (!late) { s_x[pos] = xptr; s_w[pos] = w; }
          if constexpr (DYN)
          {
            if (!late)
            {
              s_h[pos][0] = hgen[0] ? hgen[0] : p.zero;
              s_h[pos][1] = hgen[1] ? hgen[1] : p.zero;
            }
          }
          else
          {
#pragma unroll
            for (int j = 0; j < MT; ++j) s_h[pos][j] = p.term_h[size_t(tb + i) * MT + j];
          }
        }
        count += __popc(mask);
        late_count += __popc(mask_late);
        if (DYN && p.dyn.stats && blockIdx.y == 0)
        {
          // traffic accounting: H and X partitions this CTA reads
          const int nh = __popc(__ballot_sync(0xffffffffu, hgen[0] != nullptr))
              + __popc(__ballot_sync(0xffffffffu, hgen[1] != nullptr));
          if (lane == 0 && nh)
          {
            atomicAdd(p.dyn.stats, (unsigned long long)nh);
            // X[n, p] counts once: the old-state group re-reads it from L2
            if (kind != 1) atomicAdd(p.dyn.stats + 1, (unsigned long long)__popc(mask | mask_late));
          }
        }
      }
      if (lane == 0) { s_count = count; if (DYN) s_late = late_count; }
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
    if (last_pass) break;
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
  }  // segments
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
  return size_t(STAGES) * (1 + MT) * kTmaTileBytes + size_t(STAGES) * (8 + 8 + 4 + 4 + 4) + 128;
}

// Stage markers the producer hands to the consumers next to the data stages
constexpr int kStageData = 0, kStageDone = 1, kStageFlush = 2;

// Producer side, shared by the streaming and the time-batched kernel: resolves the
// 64-term tile [tb, tb + tile_n) of the term table into (x pointer(s), h pointers, weight)
// records in shared memory, dropping terms with zero weight (a muted source costs no
// traffic).  Returns the number of records.
// defer_count != null (forward transform still running, MacParams::head_bias): terms with
// p == 0 -- the only ones that read the spectrum the forward transform is producing -- go
// to the side list d_* instead (at most `defer_cap` of them; *defer_count is updated; the
// caller checks the room beforehand).
constexpr int kMacDeferCap = 64;

template<int MT, int TB>
__device__ __forceinline__ int mac_tma_resolve_tile(const MacParams& p, int tb, int tile_n, int lane, int nvec,
                                                    const float4* (*q_x)[TB], const float4* (*q_h)[MT], float* q_w,
                                                    int* defer_count = nullptr, const float4** d_x = nullptr,
                                                    const float4* (*d_h)[MT] = nullptr, float* d_w = nullptr)
{
  int count = 0;
  int dcount = defer_count ? *defer_count : 0;
  for (int base = 0; base < tile_n; base += 32)
  {
    const int i = base + lane;
    bool ok = false, late = false;
    MacTermMeta m; float w = 0.f;
    if (i < tile_n)
    {
      m = p.term_meta[tb + i];
      w = p.wtab[m.wslot + p.wtab_offset];
      ok = (w != 0.0f);
      late = ok && defer_count && m.p == 0;
    }
    const unsigned mask_late = __ballot_sync(0xffffffffu, late);
    const unsigned mask = __ballot_sync(0xffffffffu, ok && !late);
    if (ok)
    {
      const int parts = p.xparts[m.input];
      const int head = p.heads[m.input] + p.head_bias;   // slot of the LAST block of the launch
      const float4* xp[TB];
#pragma unroll
      for (int ts = 0; ts < TB; ++ts)
      {
        int slot = head - (TB - 1 - ts) - m.p;
        while (slot < 0) slot += parts;
        if (slot >= parts) slot -= parts;
        xp[ts] = p.xring[m.input] + size_t(slot) * nvec;
      }
      if (late)
      {
        const int pos = dcount + __popc(mask_late & ((1u << lane) - 1u));
        d_x[pos] = xp[0];
        d_w[pos] = w;
#pragma unroll
        for (int j = 0; j < MT; ++j) d_h[pos][j] = p.term_h[size_t(tb + i) * MT + j];
      }
      else
      {
        const int pos = count + __popc(mask & ((1u << lane) - 1u));
#pragma unroll
        for (int ts = 0; ts < TB; ++ts) q_x[pos][ts] = xp[ts];
        q_w[pos] = w;
#pragma unroll
        for (int j = 0; j < MT; ++j) q_h[pos][j] = p.term_h[size_t(tb + i) * MT + j];
      }
    }
    count += __popc(mask);
    dcount += __popc(mask_late);
  }
  if (defer_count) *defer_count = dcount;
  __syncwarp();
  return count;
}

template<int MT, int STAGES>
__global__ void __launch_bounds__(kTmaThreads, 1)
mac_tma_kernel(const MacParams p)
{
  extern __shared__ __align__(128) unsigned char tma_smem[];
  __shared__ const float4* q_x[kMacTile][1];       // producer-private staging of a term tile
  __shared__ const float4* q_h[kMacTile][MT];
  __shared__ float q_w[kMacTile];
  __shared__ const float4* d_x[kMacDeferCap];      // p == 0 terms held back until the forward transform is done
  __shared__ const float4* d_h[kMacDeferCap][MT];
  __shared__ float d_w[kMacDeferCap];
  float4* buf = reinterpret_cast<float4*>(tma_smem);
  uint64_t* full = reinterpret_cast<uint64_t*>(tma_smem + size_t(STAGES) * (1 + MT) * kTmaTileBytes);
  uint64_t* empty = full + STAGES;
  volatile float* stage_w = reinterpret_cast<volatile float*>(empty + STAGES);
  volatile int* stage_kind = reinterpret_cast<volatile int*>(const_cast<float*>(stage_w) + STAGES);
  volatile int* stage_aux = stage_kind + STAGES;   // kStageFlush: partial index the accumulators go to

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
    if (p.trace) { unsigned long long now; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(now)); p.trace[size_t(blockIdx.x) * 4] = now; }
  }
  __syncthreads();
  // head_bias: this block's forward transform may still be running (programmatic dependent
  // launch); everything it does not produce -- the term tables, the weights, the delay-line
  // slots of EARLIER blocks, the filter spectra -- may be read already.  Only the producer
  // warp reads global memory; it waits before its first p == 0 term.
  if (!p.head_bias) grid_dependency_wait();
  // the kernel behind this one (the block's inverse transform, or the next accumulator kind's MAC)
  // may move into SMs as they free up: its launch latency and prologue hide under our tail
  if (p.early_trigger) grid_launch_dependents();

  if (t >= kThreads)
  {
    // ---- producer warp: resolve the terms tile by tile, then one lane feeds the ring.
    // Start-up latency is what a 25 - 50 us launch pays for most (profiles/r2_mac_balance.md: first data
    // after 4 - 10 us): the dependent loads from blockIdx to the first bulk copy are kept to three levels
    // (CTA head -> term records -> weights / ring heads), and the records of the NEXT tile are
    // fetched before the current one is fed, so that later tiles cost one level.
    const uint64_t pol_h = p.l2_hints ? l2_policy_evict_first() : l2_policy_evict_normal();
    const uint64_t pol_x = p.l2_hints ? l2_policy_evict_last() : l2_policy_evict_normal();
    const int4 hd = p.cta_head[2 * blockIdx.x];
    const int seg_first = hd.x, seg_last = hd.y;
    int s = 0;
    unsigned phase = 0;
    bool waited = !p.head_bias;
    auto issue = [&](const float4* x, const float4* const* h, float w)
    {
      mbar_wait(empty + s, phase ^ 1u);              // the consumers have released this stage
      stage_w[s] = w;
      stage_kind[s] = kStageData;
      mbar_arrive_expect_tx(full + s, (1 + MT) * kTmaTileBytes);
      float4* dst = buf + size_t(s) * (1 + MT) * kTmaTileVec;
      bulk_copy_g2s(dst, x + tile_off, kTmaTileBytes, full + s, pol_x);
#pragma unroll
      for (int j = 0; j < MT; ++j)
        bulk_copy_g2s(dst + size_t(1 + j) * kTmaTileVec, h[j] + tile_off, kTmaTileBytes, full + s, pol_h);
      if (++s == STAGES) { s = 0; phase ^= 1u; }
    };
    bool first_issue = p.trace != nullptr && p.trace_mode == 1;
    auto feed = [&](int count)
    {
      if (lane == 0)
      {
        if (first_issue && count > 0)
        { unsigned long long now; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(now)); p.trace[size_t(blockIdx.x) * 4 + 3] = now; first_issue = false; }
        for (int i = 0; i < count; ++i) issue(q_x[i][0], q_h[i], q_w[i]);
      }
      __syncwarp();
    };
    auto marker = [&](int kind, int aux)
    {
      if (lane == 0)
      {
        mbar_wait(empty + s, phase ^ 1u);
        stage_kind[s] = kind;
        stage_aux[s] = aux;
        mbar_arrive(full + s);
        if (++s == STAGES) { s = 0; phase ^= 1u; }
      }
      __syncwarp();
    };
    // Until the forward transform of this block is known to be complete, terms with
    // p == 0 are resolved but held back in the side list; everything else streams.  The
    // wait comes when the CTA's first segment ends or the side list is full.
    int deferred = 0;
    auto release_deferred = [&]()
    {
      grid_dependency_wait();                        // the forward transform of this block is complete
      asm volatile("fence.proxy.async.global;" ::: "memory");
      waited = true;
      if (lane == 0)
        for (int i = 0; i < deferred; ++i) issue(d_x[i], d_h[i], d_w[i]);
      __syncwarp();
      deferred = 0;
    };
    // a tile's term records, two per lane (first load level of a tile)
    struct TileRegs { MacTermMeta m[2]; const float4* h[2][MT]; };
    auto fetch = [&](int tb, int tile_n, TileRegs& r)
    {
#pragma unroll
      for (int rr = 0; rr < 2; ++rr)
      {
        const int i = rr * 32 + lane;
        if (i < tile_n)
        {
          r.m[rr] = p.term_meta[tb + i];
#pragma unroll
          for (int j = 0; j < MT; ++j) r.h[rr][j] = p.term_h[size_t(tb + i) * MT + j];
        }
      }
    };
    // second level (weights, ring heads), then the records go to the staging lists; terms with
    // zero weight are dropped (a muted source costs no traffic)
    auto finish = [&](const TileRegs& r, int tile_n) -> int
    {
      float w[2]; int parts[2], head[2]; const float4* xr[2];
#pragma unroll
      for (int rr = 0; rr < 2; ++rr)
      {
        w[rr] = 0.f; parts[rr] = 1; head[rr] = 0; xr[rr] = nullptr;
        if (rr * 32 + lane < tile_n)
        {
          w[rr] = p.wtab[r.m[rr].wslot + p.wtab_offset];
          parts[rr] = p.xparts[r.m[rr].input];
          head[rr] = *reinterpret_cast<const volatile int*>(p.heads + r.m[rr].input);   // (see fwd_fft_body)
          xr[rr] = p.xring[r.m[rr].input];
        }
      }
      int count = 0;
#pragma unroll
      for (int rr = 0; rr < 2; ++rr)
      {
        const bool ok = (rr * 32 + lane < tile_n) && w[rr] != 0.0f;
        const bool late = ok && !waited && r.m[rr].p == 0;
        const unsigned mask_late = __ballot_sync(0xffffffffu, late);
        const unsigned mask = __ballot_sync(0xffffffffu, ok && !late);
        if (ok)
        {
          int slot = head[rr] + p.head_bias - r.m[rr].p;
          while (slot < 0) slot += parts[rr];
          if (slot >= parts[rr]) slot -= parts[rr];
          const float4* xp = xr[rr] + size_t(slot) * nvec;
          if (late)
          {
            const int pos = deferred + __popc(mask_late & ((1u << lane) - 1u));
            d_x[pos] = xp; d_w[pos] = w[rr];
#pragma unroll
            for (int j = 0; j < MT; ++j) d_h[pos][j] = r.h[rr][j];
          }
          else
          {
            const int pos = count + __popc(mask & ((1u << lane) - 1u));
            q_x[pos][0] = xp; q_w[pos] = w[rr];
#pragma unroll
            for (int j = 0; j < MT; ++j) q_h[pos][j] = r.h[rr][j];
          }
        }
        count += __popc(mask);
        deferred += __popc(mask_late);
      }
      __syncwarp();
      return count;
    };
    // the CTA's segments beyond the first, one per lane (a CTA rarely has more than two or three)
    int4 seg_pref = make_int4(0, 0, 0, 0);
    if (seg_first + 1 + lane < seg_last) seg_pref = p.work[seg_first + 1 + lane];
    auto load_seg = [&](int sg) -> int4
    {
      const int k = sg - seg_first - 1;
      if (k >= 32) return p.work[sg];
      return make_int4(__shfl_sync(0xffffffffu, seg_pref.x, k), __shfl_sync(0xffffffffu, seg_pref.y, k),
                       __shfl_sync(0xffffffffu, seg_pref.z, k), 0);
    };

    int seg = seg_first;
    int4 wk = p.cta_head[2 * blockIdx.x + 1];
    int tb = wk.x;
    bool have = seg < seg_last;
    TileRegs cur;
    if (have) fetch(tb, min(kMacTile, wk.y - tb), cur);
    while (have)
    {
      const int tile_n = max(0, min(kMacTile, wk.y - tb));
      if (!waited && deferred + tile_n > kMacDeferCap) release_deferred();
      const int count = finish(cur, tile_n);
      const bool seg_end = tb + kMacTile >= wk.y;
      // the next tile's records are requested before this one is fed
      int nseg = seg, ntb = tb + kMacTile;
      int4 nwk = wk;
      if (seg_end)
      {
        nseg = seg + 1;
        if (nseg < seg_last) { nwk = load_seg(nseg); ntb = nwk.x; }
      }
      const bool nhave = nseg < seg_last;
      TileRegs nxt;
      if (nhave) fetch(ntb, min(kMacTile, nwk.y - ntb), nxt);
      feed(count);
      if (seg_end)
      {
        if (!waited) release_deferred();             // the side list belongs to this segment's accumulators
        marker(kStageFlush, wk.z + p.partial_offset);
      }
      seg = nseg; wk = nwk; tb = ntb; have = nhave;
      if (have) cur = nxt;
    }
    marker(kStageDone, 0);
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
  bool first_stage = true;
  while (true)
  {
    mbar_wait(full + s, phase);
    if (p.trace && t == 0 && first_stage)
    { unsigned long long now; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(now)); p.trace[size_t(blockIdx.x) * 4 + 1] = now; }
    first_stage = false;
    const int kind = stage_kind[s];
    if (kind == kStageDone)
    {
      if (p.trace && t == 0)
      {
        unsigned long long now; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(now));
        unsigned long long* tr = p.trace + size_t(blockIdx.x) * 4;
        tr[2] = now;
        if (p.trace_mode == 0) tr[3] += now - tr[0];   // only this CTA writes its slot
      }
      break;
    }
    if (kind == kStageFlush)
    {
      // end of a segment: the accumulators go to its partial spectra
      // (bin 0 holds (DC, Nyquist), both real products, convolver.h:510-541)
      const int pidx = stage_aux[s];
      if (tt == 0)
      {
#pragma unroll
        for (int j = 0; j < MT; ++j) { acc[j][0].x += nyq[j]; acc[j][0].y = nyq[j]; }
      }
      float4* out = p.partials + size_t(pidx) * MT * nvec;
#pragma unroll
      for (int j = 0; j < MT; ++j)
      {
#pragma unroll
        for (int v = 0; v < 2; ++v)
        {
          out[size_t(j) * nvec + tt + v * kThreads] = acc[j][v];
          acc[j][v] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
        nyq[j] = 0.0f;
      }
    }
    else
    {
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
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(empty + s);            // this warp is done with the stage
    if (++s == STAGES) { s = 0; phase ^= 1u; }
  }
}

// ---------------------------------------------------------------- MAC, generated terms, bulk-copy staged
// Binaural / BRS at B >= 1024 (round 2; replaces mac_kernel<2, NV, 2, 2, true> there, which stays for
// smaller blocks).  Same producer / consumer ring as mac_tma_kernel, but the work is cut by
// (source, partition) SLOT, independent of the block's control data: CTA c owns slots
// [T c / G, T (c + 1) / G) of T = sources x Pmax.  A slot is evaluated for every accumulator kind its
// source takes part in this block -- constant, or old and / or new around a filter switch or weight change
// (src/brsrenderer.h:141-210, src/binauralrenderer.h:316-343) -- from ONE copy of X[n, p]: a stage holds
// the X tile and up to four H tiles (old L/R, new L/R), the consumers keep 3 kinds x 2 ears of accumulators.
// What that buys over the kind-major split of mac_kernel<DYN>:
//   * the path from blockIdx to the first bulk copy is three load levels (control block + per-source
//     weights -> filter-ring entries + delay-line heads -> zero flags, usually absent) instead of five,
//     resolved by 32 lanes at once, and 160 KiB stay in flight per SM whatever the consumers do;
//   * X[n, p] is read once per slot, not once per kind;
//   * every CTA writes ONE set of partial spectra [CTA][kind][ear][B], which the cluster variant of the
//     inverse kernel sums itself (no separate reduction launch).
struct DynStage { float wa, wb; int ka; unsigned mask; };   // mask: bit j = H tile j present (0,1: kind ka; 2,3: new)
constexpr int kDynTiles = 5;                                // X + 4 H tiles per stage
constexpr int kDynRound = 32;                               // slots resolved per round (one per producer lane)

template<int STAGES>
constexpr size_t mac_dyn_tma_smem_bytes()
{
  return size_t(STAGES) * kDynTiles * kTmaTileBytes + size_t(STAGES) * (8 + 8 + 4 + sizeof(DynStage)) + 128;
}

struct DynSlot { const float4* x; const float4* h[4]; DynStage st; };

template<int STAGES>
__global__ void __launch_bounds__(kTmaThreads, 1)
mac_dyn_tma_kernel(const MacParams p)
{
  extern __shared__ __align__(128) unsigned char tma_smem[];
  __shared__ DynSlot q[kDynRound];               // producer-private staging of a round of slots
  __shared__ DynSlot dq[kMacDeferCap];           // p == 0 slots held back until the forward transform is done
  float4* buf = reinterpret_cast<float4*>(tma_smem);
  uint64_t* full = reinterpret_cast<uint64_t*>(tma_smem + size_t(STAGES) * kDynTiles * kTmaTileBytes);
  uint64_t* empty = full + STAGES;
  volatile int* stage_kind = reinterpret_cast<volatile int*>(empty + STAGES);
  volatile DynStage* stage_info = reinterpret_cast<volatile DynStage*>(const_cast<int*>(stage_kind) + STAGES);

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
    if (p.trace) { unsigned long long now; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(now)); p.trace[size_t(blockIdx.x) * 4] = now; }
  }
  __syncthreads();
  if (!p.head_bias) grid_dependency_wait();
  if (p.early_trigger) grid_launch_dependents();   // the inverse kernel's CTAs may move in as SMs free up

  if (t >= kThreads)
  {
    const uint64_t pol_h = p.l2_hints ? l2_policy_evict_first() : l2_policy_evict_normal();
    const uint64_t pol_x = p.l2_hints ? l2_policy_evict_last() : l2_policy_evict_normal();
    int s = 0;
    unsigned phase = 0;
    bool waited = !p.head_bias;
    auto issue = [&](const DynSlot& sl)
    {
      mbar_wait(empty + s, phase ^ 1u);
      stage_info[s].wa = sl.st.wa; stage_info[s].wb = sl.st.wb;
      stage_info[s].ka = sl.st.ka; stage_info[s].mask = sl.st.mask;
      stage_kind[s] = kStageData;
      mbar_arrive_expect_tx(full + s, unsigned(1 + __popc(sl.st.mask)) * kTmaTileBytes);
      float4* dst = buf + size_t(s) * kDynTiles * kTmaTileVec;
      bulk_copy_g2s(dst, sl.x + tile_off, kTmaTileBytes, full + s, pol_x);
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (sl.st.mask & (1u << j))
          bulk_copy_g2s(dst + size_t(1 + j) * kTmaTileVec, sl.h[j] + tile_off, kTmaTileBytes, full + s, pol_h);
      if (++s == STAGES) { s = 0; phase ^= 1u; }
    };
    int deferred = 0;
    auto release_deferred = [&]()
    {
      grid_dependency_wait();                        // the forward transform of this block is complete
      asm volatile("fence.proxy.async.global;" ::: "memory");
      waited = true;
      if (lane == 0)
        for (int i = 0; i < deferred; ++i) issue(dq[i]);
      __syncwarp();
      deferred = 0;
    };

    bool first_issue = p.trace != nullptr;
    const DynTables& d = p.dyn;
    const long long T = (long long)d.nsrc * d.Pmax;
    const int s0 = int(T * blockIdx.x / gridDim.x), s1 = int(T * (blockIdx.x + 1) / gridDim.x);
    unsigned long long n_h = 0, n_x = 0;
    for (int base = s0; base < s1; base += kDynRound)
    {
      const int i = base + lane;
      const bool in = i < s1;
      const int n = in ? i / d.Pmax : 0;
      const int pp = in ? i % d.Pmax : 0;
      // level 1: the block's control word and everything indexed by the source alone
      const long long b = d.ctl->b;
      float w0 = 0.f, w1 = 0.f, w2 = 0.f;
      int input = 0, sparts = 0;
      if (in)
      {
        w0 = p.wtab[n]; w1 = p.wtab[d.nsrc + n]; w2 = p.wtab[2 * d.nsrc + n];
        input = d.src_input[n]; sparts = d.src_parts[n];
      }
      bool alive = in && pp < sparts && (w0 != 0.0f || w1 != 0.0f || w2 != 0.0f);
      // constant sources carry kind 0 alone; the others kind 1 (old state) and / or kind 2 (new state)
      DynSlot sl;
      sl.st.ka = w0 != 0.0f ? 0 : (w1 != 0.0f ? 1 : 2);
      sl.st.wa = w0 != 0.0f ? w0 : (w1 != 0.0f ? w1 : w2);
      const bool second = w0 == 0.0f && w1 != 0.0f && w2 != 0.0f;
      sl.st.wb = second ? w2 : 0.0f;
      sl.st.mask = 0;
      // level 2: the filters current in block b - p (state after this block's switch) / b - 1 - p
      // (state before it), per ear; and the delay-line slot
      DynRingEntry en[4];
      int xparts = 1, head = 0;
      const float4* xr = nullptr;
      if (alive)
      {
#pragma unroll
        for (int k2 = 0; k2 < 2; ++k2)
        {
          const int kind = k2 == 0 ? sl.st.ka : 2;
          const bool want = k2 == 0 || second;
          long long time = b - (kind == 1 ? 1 : 0) - pp;
          int slot = int(time % d.R);
          if (slot < 0) slot += d.R;
          const bool current = (kind != 1 && pp == 0);   // ring[b] = this block's update
#pragma unroll
          for (int ear = 0; ear < 2; ++ear)
          {
            if (want)
              en[k2 * 2 + ear] = current ? d.upd[size_t(n) * 2 + ear]
                                         : d.ring[(size_t(n) * 2 + ear) * d.R + slot];
            else en[k2 * 2 + ear] = DynRingEntry{nullptr, nullptr, 0, 0};
          }
        }
        xparts = p.xparts[input]; head = *reinterpret_cast<const volatile int*>(p.heads + input); xr = p.xring[input];   // (see fwd_fft_body)
      }
      // level 3: zero flags (null for filters without zero-flagged partitions)
      if (alive)
      {
#pragma unroll
        for (int j = 0; j < 4; ++j)
        {
          sl.h[j] = nullptr;
          if (en[j].base && pp < en[j].nparts && (!en[j].flags || en[j].flags[pp]))
          {
            sl.h[j] = en[j].base + size_t(pp) * nvec;
            sl.st.mask |= 1u << j;
          }
        }
        alive = sl.st.mask != 0;
        int slot = head + p.head_bias - pp;
        while (slot < 0) slot += xparts;
        if (slot >= xparts) slot -= xparts;
        sl.x = xr + size_t(slot) * nvec;
      }
      const bool late = alive && !waited && pp == 0;
      const unsigned m_now = __ballot_sync(0xffffffffu, alive && !late);
      const unsigned m_late = __ballot_sync(0xffffffffu, late);
      if (!waited && deferred + __popc(m_late) > kMacDeferCap) release_deferred();   // (then these go out behind them)
      if (alive)
      {
        if (late && !waited) dq[deferred + __popc(m_late & ((1u << lane) - 1u))] = sl;
        else q[__popc((waited ? (m_now | m_late) : m_now) & ((1u << lane) - 1u))] = sl;
        n_h += __popc(sl.st.mask); n_x += 1;
      }
      const int count = waited ? __popc(m_now | m_late) : __popc(m_now);
      if (!waited) deferred += __popc(m_late);
      __syncwarp();
      if (lane == 0)
      {
        if (first_issue && count > 0)
        { unsigned long long now; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(now)); p.trace[size_t(blockIdx.x) * 4 + 3] = now; first_issue = false; }
        for (int k = 0; k < count; ++k) issue(q[k]);
      }
      __syncwarp();
    }
    if (!waited) release_deferred();     // (every CTA waits once: MAC complete implies forward transform complete)
    if (lane == 0)
    {
      mbar_wait(empty + s, phase ^ 1u);
      stage_kind[s] = kStageDone;
      mbar_arrive(full + s);
    }
    if (d.stats && blockIdx.y == 0)
    {
      // traffic accounting: H and X partitions this CTA read
#pragma unroll
      for (int o = 16; o > 0; o >>= 1)
      {
        n_h += __shfl_xor_sync(0xffffffffu, n_h, o);
        n_x += __shfl_xor_sync(0xffffffffu, n_x, o);
      }
      if (lane == 0 && n_h) { atomicAdd(d.stats, n_h); atomicAdd(d.stats + 1, n_x); }
    }
    return;
  }

  // ---- consumer warps: acc[kind][ear][2 float4]
  float4 acc[3][2][2];
  float nyq[3][2];
#pragma unroll
  for (int k = 0; k < 3; ++k)
#pragma unroll
    for (int j = 0; j < 2; ++j)
    {
      nyq[k][j] = 0.0f;
      acc[k][j][0] = make_float4(0.f, 0.f, 0.f, 0.f);
      acc[k][j][1] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
  const int tt = tile_off + t;
  int s = 0;
  unsigned phase = 0;
  // (kinds without sources this block are not written; read now, not in the tail)
  const int n0 = p.dyn.ctl->n[0], n1 = p.dyn.ctl->n[1], n2 = p.dyn.ctl->n[2];

#define SSR_DYN_CMAC(A, NQ, X, H) \
  { \
    A.x = fmaf(X.x, H.x, A.x); A.x = fmaf(-X.y, H.y, A.x); \
    A.y = fmaf(X.x, H.y, A.y); A.y = fmaf(X.y, H.x, A.y); \
    A.z = fmaf(X.z, H.z, A.z); A.z = fmaf(-X.w, H.w, A.z); \
    A.w = fmaf(X.z, H.w, A.w); A.w = fmaf(X.w, H.z, A.w); \
    if (v == 0 && tt == 0) NQ = fmaf(X.y, H.y, NQ); \
  }
#define SSR_DYN_KIND(K, X, ROW0, BIT0) \
  { \
    _Pragma("unroll") for (int j = 0; j < 2; ++j) \
      if (mask & (1u << (BIT0 + j))) \
      { \
        const float4 h = sb[size_t(ROW0 + j) * kTmaTileVec + t + v * kThreads]; \
        SSR_DYN_CMAC(acc[K][j][v], nyq[K][j], X, h) \
      } \
  }

  bool first_stage = p.trace != nullptr && t == 0;
  while (true)
  {
    mbar_wait(full + s, phase);
    if (first_stage)
    { unsigned long long now; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(now)); p.trace[size_t(blockIdx.x) * 4 + 1] = now; first_stage = false; }
    if (stage_kind[s] == kStageDone) break;
    const float wa = stage_info[s].wa, wb = stage_info[s].wb;
    const int ka = stage_info[s].ka;
    const unsigned mask = stage_info[s].mask;
    const float4* sb = buf + size_t(s) * kDynTiles * kTmaTileVec;
#pragma unroll
    for (int v = 0; v < 2; ++v)
    {
      const float4 x = sb[t + v * kThreads];
      const float4 xa = make_float4(x.x * wa, x.y * wa, x.z * wa, x.w * wa);
      if (ka == 0) SSR_DYN_KIND(0, xa, 1, 0)
      else if (ka == 1) SSR_DYN_KIND(1, xa, 1, 0)
      else SSR_DYN_KIND(2, xa, 1, 0)
      if (mask & 12u)
      {
        const float4 xb = make_float4(x.x * wb, x.y * wb, x.z * wb, x.w * wb);
        SSR_DYN_KIND(2, xb, 3, 2)
      }
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(empty + s);
    if (++s == STAGES) { s = 0; phase ^= 1u; }
  }
#undef SSR_DYN_KIND
#undef SSR_DYN_CMAC

  // one set of partial spectra per CTA: [CTA][kind][ear][B]; kinds without sources this block are
  // skipped (the inverse kernel skips them too); bin 0 holds (DC, Nyquist), both real products
  // (convolver.h:510-541)
#pragma unroll
  for (int k = 0; k < 3; ++k)
  {
    if ((k == 0 ? n0 : (k == 1 ? n1 : n2)) == 0) continue;
#pragma unroll
    for (int j = 0; j < 2; ++j)
    {
      if (tt == 0) { acc[k][j][0].x += nyq[k][j]; acc[k][j][0].y = nyq[k][j]; }
      float4* out = p.partials + ((size_t(blockIdx.x) * 3 + k) * 2 + j) * nvec;
      out[tt] = acc[k][j][0];
      out[tt + kThreads] = acc[k][j][1];
    }
  }
  if (p.trace && t == 0)
  { unsigned long long now; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(now)); p.trace[size_t(blockIdx.x) * 4 + 2] = now; }
}

// ---------------------------------------------------------------- MAC, time-batched (offline)
// Offline rendering knows every input block in advance (apf::mimoprocessor_file_io).
// TB consecutive blocks are then rendered per launch: every H tile is read ONCE and
// multiplied with the TB delay-line spectra X[n, p] of the TB time steps, so the
// H stream -- the whole cost of the streaming MAC -- is amortised over TB blocks
// ("an offline batched mode may be reported separately"; it is not the
// real-time path, where block t + 1 does not exist yet when block t is due).
// Same producer / consumer structure as mac_tma_kernel; a stage holds the MT H tiles
// and the TB X tiles of one term.  For time step t of the batch the term (n, p)
// reads the delay-line slot of block (last - (TB - 1 - t)) - p, i.e. TB consecutive
// slots; the ring therefore has TB - 1 spare slots (Input::slots).
// Partials: [CTA][t][row][B].
template<int MT, int TB, int STAGES>
constexpr size_t mac_tma_batch_smem_bytes()
{
  return size_t(STAGES) * (TB + MT) * kTmaTileBytes + size_t(STAGES) * (8 + 8 + 4 + 4 + 4) + 128;
}

template<int MT, int TB, int STAGES>
__global__ void __launch_bounds__(kTmaThreads, 1)
mac_tma_batch_kernel(const MacParams p)
{
  extern __shared__ __align__(128) unsigned char tma_smem[];
  __shared__ const float4* q_x[kMacTile][TB];
  __shared__ const float4* q_h[kMacTile][MT];
  __shared__ float q_w[kMacTile];
  constexpr int kStageVec = (TB + MT) * kTmaTileVec;
  float4* buf = reinterpret_cast<float4*>(tma_smem);
  uint64_t* full = reinterpret_cast<uint64_t*>(tma_smem + size_t(STAGES) * (TB + MT) * kTmaTileBytes);
  uint64_t* empty = full + STAGES;
  volatile float* stage_w = reinterpret_cast<volatile float*>(empty + STAGES);
  volatile int* stage_kind = reinterpret_cast<volatile int*>(const_cast<float*>(stage_w) + STAGES);
  volatile int* stage_aux = stage_kind + STAGES;

  const int t = threadIdx.x;
  const int lane = t & 31;
  const int nvec = p.B >> 1;
  const int tile_off = blockIdx.y * kTmaTileVec;

  if (t == 0)
  {
#pragma unroll
    for (int s = 0; s < STAGES; ++s) { mbar_init(full + s, 1); mbar_init(empty + s, kThreads / 32); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncthreads();
  grid_dependency_wait();
  const int seg_first = p.cta_first[blockIdx.x], seg_last = p.cta_first[blockIdx.x + 1];

  if (t >= kThreads)
  {
    const uint64_t pol_h = p.l2_hints ? l2_policy_evict_first() : l2_policy_evict_normal();
    const uint64_t pol_x = p.l2_hints ? l2_policy_evict_last() : l2_policy_evict_normal();
    int s = 0;
    unsigned phase = 0;
    for (int seg = seg_first; seg <= seg_last; ++seg)
    {
      int4 wk = make_int4(0, 0, 0, 0);
      if (seg < seg_last)
      {
        wk = p.work[seg];
        for (int tb = wk.x; tb < wk.y; tb += kMacTile)
        {
          const int count = mac_tma_resolve_tile<MT, TB>(p, tb, min(kMacTile, wk.y - tb), lane, nvec, q_x, q_h, q_w);
          if (lane == 0)
          {
            for (int i = 0; i < count; ++i)
            {
              mbar_wait(empty + s, phase ^ 1u);
              stage_w[s] = q_w[i];
              stage_kind[s] = kStageData;
              mbar_arrive_expect_tx(full + s, (TB + MT) * kTmaTileBytes);
              float4* dst = buf + size_t(s) * kStageVec;
#pragma unroll
              for (int ts = 0; ts < TB; ++ts)
                bulk_copy_g2s(dst + size_t(ts) * kTmaTileVec, q_x[i][ts] + tile_off, kTmaTileBytes, full + s, pol_x);
#pragma unroll
              for (int j = 0; j < MT; ++j)
                bulk_copy_g2s(dst + size_t(TB + j) * kTmaTileVec, q_h[i][j] + tile_off, kTmaTileBytes, full + s, pol_h);
              if (++s == STAGES) { s = 0; phase ^= 1u; }
            }
          }
          __syncwarp();
        }
      }
      if (lane == 0)
      {
        // end of a segment (flush the accumulators to its partials) / end of the work
        mbar_wait(empty + s, phase ^ 1u);
        stage_kind[s] = seg < seg_last ? kStageFlush : kStageDone;
        stage_aux[s] = wk.z + p.partial_offset;
        mbar_arrive(full + s);
        if (++s == STAGES) { s = 0; phase ^= 1u; }
      }
      __syncwarp();
    }
    return;
  }

  // ---- consumers: acc[time step][row][2 float4]
  float4 acc[TB][MT][2];
  float nyq[TB][MT];
#pragma unroll
  for (int ts = 0; ts < TB; ++ts)
#pragma unroll
    for (int j = 0; j < MT; ++j)
    {
      nyq[ts][j] = 0.0f;
      acc[ts][j][0] = make_float4(0.f, 0.f, 0.f, 0.f);
      acc[ts][j][1] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
  const int tt = tile_off + t;
  int s = 0;
  unsigned phase = 0;
  while (true)
  {
    mbar_wait(full + s, phase);
    const int kind = stage_kind[s];
    if (kind == kStageDone) break;
    if (kind == kStageFlush)
    {
      const int pidx = stage_aux[s];
      float4* out = p.partials + size_t(pidx) * TB * MT * nvec;   // partials: [segment][t][row][B]
#pragma unroll
      for (int ts = 0; ts < TB; ++ts)
#pragma unroll
        for (int j = 0; j < MT; ++j)
        {
          if (tt == 0) { acc[ts][j][0].x += nyq[ts][j]; acc[ts][j][0].y = nyq[ts][j]; }
#pragma unroll
          for (int v = 0; v < 2; ++v)
          {
            out[(size_t(ts) * MT + j) * nvec + tt + v * kThreads] = acc[ts][j][v];
            acc[ts][j][v] = make_float4(0.f, 0.f, 0.f, 0.f);
          }
          nyq[ts][j] = 0.0f;
        }
    }
    else
    {
      const float w = stage_w[s];
      const float4* sb = buf + size_t(s) * kStageVec;
#pragma unroll
      for (int v = 0; v < 2; ++v)
      {
        float4 h[MT];
#pragma unroll
        for (int j = 0; j < MT; ++j) h[j] = sb[size_t(TB + j) * kTmaTileVec + t + v * kThreads];
#pragma unroll
        for (int ts = 0; ts < TB; ++ts)
        {
          float4 x = sb[size_t(ts) * kTmaTileVec + t + v * kThreads];
          x.x *= w; x.y *= w; x.z *= w; x.w *= w;
#pragma unroll
          for (int j = 0; j < MT; ++j)
          {
            acc[ts][j][v].x = fmaf(x.x, h[j].x, acc[ts][j][v].x);
            acc[ts][j][v].x = fmaf(-x.y, h[j].y, acc[ts][j][v].x);
            acc[ts][j][v].y = fmaf(x.x, h[j].y, acc[ts][j][v].y);
            acc[ts][j][v].y = fmaf(x.y, h[j].x, acc[ts][j][v].y);
            acc[ts][j][v].z = fmaf(x.z, h[j].z, acc[ts][j][v].z);
            acc[ts][j][v].z = fmaf(-x.w, h[j].w, acc[ts][j][v].z);
            acc[ts][j][v].w = fmaf(x.z, h[j].w, acc[ts][j][v].w);
            acc[ts][j][v].w = fmaf(x.w, h[j].z, acc[ts][j][v].w);
            if (v == 0 && tt == 0) nyq[ts][j] = fmaf(x.y, h[j].y, nyq[ts][j]);
          }
        }
      }
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(empty + s);
    if (++s == STAGES) { s = 0; phase ^= 1u; }
  }
}

}  // namespace ssrk
