// inverse.cuh -- partial reduction, inverse transform + crossfade + source sum, the fused multi-GPU output gather and the filter-ring update
// (part of the kernel set of kernels.cuh; included from there, in this order)
#pragma once

#include <cooperative_groups.h>

namespace ssrk
{

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

constexpr int kInvMaxRes = 8;        // B <= 4096: (B/2 float2) / 256 threads
constexpr int kInvMaxResLarge = 16;  // B <= 8192 (inv_fft_large_kernel)

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
  unsigned long long timeout_ns;           // 0: kGatherTimeoutNs
  // host-direct delivery (HostGather): InvJob::out points into a device-mapped host
  // segment; the CTA that finishes last writes flag_value to host_flag
  unsigned* cta_counter;                   // device-local, zero between launches
  unsigned cta_total;
  unsigned long long* host_flag;           // mapped host memory, or null
  unsigned long long flag_value;
};

// Ring heads left untouched by a forward launch that the MAC overlapped
// (fwd_fft_overlap_kernel): CTA 0 of the block's inverse kernel advances them once the
// MAC -- their last reader -- is complete.
struct HeadCommit
{
  const FwdJob* jobs;   // the forward launch's job table (head pointer + ring length per input)
  int n;                // 0: nothing to commit
  int early_trigger;    // 1: the kernel lets its dependents (the next block's forward transform, which waits for
                        //  this kernel's completion itself) be scheduled at once (SSR_B200_INV_TRIGGER=1, experiment)
                        // 2: BLOCK OVERLAP (Engine::block_overlap): an extra CTA commits the heads FIRST -- the MAC,
                        //  their last reader, is complete when this kernel starts -- and only then has the whole
                        //  grid let its dependents in; the next block's forward transform (late_wait) passes the trigger on at
                        //  once, so the next block's MAC streams its p != 0 terms while this kernel still runs
};

// Start of every inverse kernel of the block loop (before anything that waits); true: this CTA is done.
// early_trigger == 2: the launch carries ONE EXTRA CTA (the last one) that only commits the heads, so that the
// dependent load -> store -> fence (1.5 - 2 us) is on no output block's path.
__device__ __forceinline__ bool inv_prologue(const HeadCommit& hc, int nthreads)
{
  if (hc.early_trigger == 2)
  {
    if (blockIdx.x == gridDim.x - 1)
    {
      for (int i = threadIdx.x; i < hc.n; i += nthreads)
      {
        int* head = hc.jobs[i].head;
        int h = *head + 1;
        if (h >= hc.jobs[i].parts) h = 0;
        *head = h;
      }
      __threadfence();     // the new heads are at device scope before this CTA's trigger
      __syncthreads();
      grid_launch_dependents();
      return true;
    }
    grid_launch_dependents();
  }
  else if (hc.early_trigger) grid_launch_dependents();
  return false;
}

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

// spins until *p >= target; gives up (sticky error flag, false) after kGatherTimeoutNs
// so that a dead peer can never hang the GPU.  `error` lives in mapped host memory:
// the host reads it after the block's completion event (Renderer::wait) and fails
// the call with SSR_ERR_STATE instead of delivering a stale block.
__device__ __noinline__ bool gather_wait_ge(const unsigned long long* p, unsigned long long target, int* error,
                                            unsigned long long timeout_ns)
{
  if (timeout_ns == 0) timeout_ns = kGatherTimeoutNs;
  if (ld_acquire_sys(p) >= target) return true;
  if (error && *reinterpret_cast<volatile int*>(error)) return false;
  const unsigned long long t0 = globaltimer_ns();
  while (ld_acquire_sys(p) < target)
  {
    __nanosleep(200);
    if (globaltimer_ns() - t0 > timeout_ns)
    {
      if (error) { *reinterpret_cast<volatile int*>(error) = 1; __threadfence_system(); }
      return false;
    }
  }
  return true;
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
  unsigned long long timeout_ns;
};
constexpr int kGatherCounterStride = 16;  // unsigned long longs between counters (one 128-byte line each)

__global__ void gather_wait_kernel(GatherWaitArgs a)
{
  // (programmatic dependent launch: the root's own inverse kernel -- whose blocks go to the
  // same slot -- must be complete when this kernel is)
  grid_launch_dependents();   // the kernels behind (release, the next block's chain) wait for our completion themselves
  grid_dependency_wait();
  const int g = threadIdx.x;
  if (g < a.world && a.ctas[g] > 0)
    gather_wait_ge(a.arrived + size_t(g) * kGatherCounterStride,
        (a.epoch + 1) * (unsigned long long)a.ctas[g], a.error, a.timeout_ns);
}

// root: block `epoch` has been read out of its slot -> peers may overwrite it
__global__ void gather_release_kernel(unsigned long long* consumed, unsigned long long value)
{
  grid_launch_dependents();
  grid_dependency_wait();   // everything queued before (the readers of the slot) is complete
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

// Binaural / BRS: this block's control data (DynCtl + lists + weights + filter updates, a few KB)
// goes from the renderer's page-locked staging buffer to the device with a KERNEL instead of a copy
// in front of the block: as the first link of the block's programmatic-dependent-launch chain it is
// scheduled while the previous block's inverse transform still runs (which triggers its dependents at
// once, InvFlags::early_trigger), reads the host buffer over PCIe meanwhile, and only then waits for
// that inverse kernel -- so neither the transfer nor a launch latency sits between two blocks (a
// stream-ordered cudaMemcpyAsync costs ~5 us there, 10 % of a BRS block).  Two device buffers
// alternate by block parity (the previous block's kernels still read theirs).  `ack`: mapped host
// word that receives the number of uploads done from this staging slot, so that the host knows when
// it may refill it without an event in the stream.
// late_trigger (block overlap of the Binaural / BRS chain, Engine::block_overlap): the kernel sits BEHIND the
// block's forward transform (which no longer needs the control data: the filter-ring update moved here) and in
// front of the MAC; it lets the MAC in as soon as the copy is complete and fenced, so the MAC resolves its slots
// and streams its p != 0 terms while this kernel finishes and the forward transform -- and, in a loop of
// device-resident blocks, the previous block's inverse transform -- still run.  Its wait at the end makes its
// completion imply the forward kernel's, which implies the previous block's.
__global__ void __launch_bounds__(kThreads)
ctl_upload_kernel(const uint4* __restrict__ src_host, uint4* __restrict__ dst, int n16,
                  unsigned long long* counter, volatile unsigned long long* ack,
                  const FwdRingUpdate ru, int ring_rows, int late_trigger)
{
  if (!late_trigger) grid_launch_dependents();
  // (the upload count is fetched while the copy's loads are in flight: this kernel's whole duration is on the path
  //  of a synchronous block, whose forward transform waits for its completion)
  unsigned long long count = 0;
  if (threadIdx.x == 0) count = *counter + 1;
  for (int i = threadIdx.x; i < n16; i += kThreads)
  {
    uint4 v;
    asm volatile("ld.global.cv.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(src_host + i));
    dst[i] = v;
  }
  if (late_trigger)
  {
    __threadfence();      // the device copy is at device scope before the MAC is let in
    __syncthreads();
    grid_launch_dependents();
    // ring[(source, ear)][b mod R] = the filters current in this block (read by LATER blocks only; the MAC of
    // this block takes them from the control data; the last reader of that ring slot was the previous block's MAC)
    if (ru.ctl)
      for (int row = threadIdx.x; row < ring_rows; row += kThreads) fwd_ring_update(ru, row >> 1, row & 1);
  }
  __syncthreads();
  if (threadIdx.x == 0)
  {
    // every load of the staging slot has returned: the host may refill it.  Upload-first chain (a synchronous
    // block: this kernel's duration is on its path, and the kernel ends at once): a posted write, no fence.
    // Late-trigger chain (a loop of queued blocks): the kernel now lingers until the forward transform and the
    // previous block are through, and the host -- two blocks ahead -- is waiting for exactly this word.
    *counter = count;
    *ack = count;
    if (late_trigger) __threadfence_system();
  }
  grid_dependency_wait();   // completion of this kernel implies completion of everything in front of it
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

// CLUSTER (generated-terms MAC, mac_dyn_tma_kernel): one thread-block cluster per output block.  The
// partial spectra [CTA][kind][ear][B] of the MAC launch are summed by the whole cluster -- CTA r takes
// the r-th slice of the bins, group g the accumulator kind g -- straight into the leader CTA's
// shared memory over DSMEM; the leader then transforms as usual.  This replaces the separate
// dyn_reduce_kernel launch (9 us of a 47 us BRS block).
// MAXRES: output pairs per thread (B / 2 / kThreads); TWS: twiddle table staged in shared memory (B <= 4096;
// above that the transform buffers take the shared memory and the twiddles come from L1 / L2)
template<int GROUPS, bool CLUSTER, int MAXRES = kInvMaxRes, bool TWS = true>
__device__ __forceinline__ void
inv_fft_body(const InvJob* __restrict__ jobs, const InvEntry* __restrict__ entries,
             const float2* __restrict__ partials, const float2* __restrict__ reduced, int B,
             const float2* __restrict__ tw,
             const float* __restrict__ fade_out, const float* __restrict__ fade_in,
             const GatherArgs& ga, const int* __restrict__ kind_counts, const HeadCommit& hc, int dyn_G,
             unsigned long long* dbg = nullptr)
{
  // (development: per-CTA time stamps of the cluster variant, scripts/mac_trace2.py)
#define SSR_INV_STAMP(i) \
  if constexpr (CLUSTER) { if (dbg && threadIdx.x == 0) dbg[size_t(blockIdx.x) * 8 + (i)] = globaltimer_ns(); }
  SSR_INV_STAMP(0)
  extern __shared__ float2 smem[];
  __shared__ float s_max[kThreads / 32];
  __shared__ int s_slot_ok;
  const int g = GROUPS == 1 ? 0 : threadIdx.x / kThreads;  // entry group
  const int t = threadIdx.x - g * kThreads;                // thread within the group
  constexpr int ngroups = GROUPS;
  float2* a = smem + size_t(g) * 2 * B;
  float2* b = a + B;
  unsigned crank = 0, csize = 1;
  if constexpr (CLUSTER)
  {
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(crank));
    asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(csize));
  }
  const InvJob job = jobs[CLUSTER ? blockIdx.x / csize : blockIdx.x];
  const int half = B >> 1;
  // twiddle table -> shared memory, once for all groups
  float2* stw = smem + size_t(GROUPS) * 2 * B;
  if (TWS && crank == 0)
    stage_twiddles(stw, tw, B, threadIdx.x, kThreads * GROUPS);  // constants: may overlap the predecessor's tail
  // (CLUSTER: static tables and the block's control word -- uploaded before the block's first kernel --
  // are read ahead of the wait; they are the head of the dependent chain to the partial spectra)
  InvEntry cen{};
  int ckind_count = 0;
  if constexpr (CLUSTER)
  {
    if (g < job.n_entries)
    {
      cen = entries[job.first_entry + g];
      ckind_count = kind_counts[cen.fade];
    }
  }
  grid_dependency_wait();
  SSR_INV_STAMP(1)
  if constexpr (CLUSTER)
  {
    namespace cg = cooperative_groups;
    cg::cluster_group cluster = cg::this_cluster();
    cluster.sync();   // every CTA of the cluster is running: the leader's shared memory may be written
    SSR_INV_STAMP(2)
    if (g < job.n_entries)
    {
      const InvEntry en = cen;
      if (ckind_count != 0)
      {
        const int nvec = B >> 1;
        const int nslice = nvec / int(csize);   // float4 per CTA of the cluster
        float4* dst = reinterpret_cast<float4*>(cluster.map_shared_rank(smem + size_t(g) * 2 * B, 0))
            + size_t(crank) * nslice;
        const float4* base = reinterpret_cast<const float4*>(partials)
            + (size_t(en.fade) * 2 + en.part_row) * nvec + size_t(crank) * nslice;
        const size_t cstride = size_t(6) * nvec;   // partials: [CTA][kind][ear][B]
        // A warp reads one CTA's whole slice row (32 lanes x float4 = 512 contiguous bytes) per load;
        // the 8 warps of the group split the CTAs, each with its loads of a batch all in flight before
        // the first add (the phase is L2 latency: two round trips for 148 partial sets); the warps'
        // sums meet in the group's spare transform buffer.  Fixed order: deterministic.
        constexpr int UNR = 10, NW = kThreads / 32;
        const int warp = t >> 5, lane = t & 31;
        float4* scratch = reinterpret_cast<float4*>(b);
        for (int r0 = 0; r0 < nslice; r0 += 32)
        {
          const int v = r0 + lane;
          float4 s0 = make_float4(0.f, 0.f, 0.f, 0.f);
          for (int c0 = warp; c0 < dyn_G; c0 += NW * UNR)
          {
            float4 q[UNR];
#pragma unroll
            for (int u = 0; u < UNR; ++u)
            {
              const int c = c0 + u * NW;
              q[u] = (v < nslice && c < dyn_G) ? __ldcs(base + size_t(c) * cstride + v) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
#pragma unroll
            for (int u = 0; u < UNR; ++u) { s0.x += q[u].x; s0.y += q[u].y; s0.z += q[u].z; s0.w += q[u].w; }
          }
          scratch[warp * 32 + lane] = s0;
          group_sync(g + 1);
          if (warp == 0)
          {
#pragma unroll
            for (int w2 = 1; w2 < NW; ++w2)
            {
              const float4 o = scratch[w2 * 32 + lane];
              s0.x += o.x; s0.y += o.y; s0.z += o.z; s0.w += o.w;
            }
            if (v < nslice) dst[v] = s0;
          }
          group_sync(g + 1);
        }
      }
    }
    SSR_INV_STAMP(3)
    cluster.sync();   // (release / acquire: the slices are visible in the leader's shared memory)
    SSR_INV_STAMP(4)
    if (crank != 0) return;
  }
  if (hc.n && hc.early_trigger != 2 && blockIdx.x == 0)
    for (int i = threadIdx.x; i < hc.n; i += kThreads * GROUPS)
    {
      int* head = hc.jobs[i].head;
      int h = *head + 1;
      if (h >= hc.jobs[i].parts) h = 0;
      *head = h;
    }
  __syncthreads();
  // flow control of the fused gather: the slot this block goes to must have been
  // read out by the root (checked by one thread while the others start loading)
  if (ga.consumed && threadIdx.x == 0) s_slot_ok = gather_wait_ge(ga.consumed, ga.need_consumed, ga.error, ga.timeout_ns) ? 1 : 0;

  float2 res[MAXRES];
#pragma unroll
  for (int q = 0; q < MAXRES; ++q) res[q] = make_float2(0.f, 0.f);

  // entries beyond the launched groups are handled by the last group in turn
  for (int e = g; e < job.n_entries; e += ngroups)
  {
    const InvEntry en = entries[job.first_entry + e];
    // dynamic renderers: accumulator kinds without sources this block are skipped
    if (kind_counts && kind_counts[en.fade] == 0) continue;
    // A[k] = sum over split partials
    // (float4 = two bins per thread; the loop over partials is unrolled so that
    // 8 independent loads are in flight per thread)
    if constexpr (CLUSTER)
    {
      // a[] already holds the sum (written by the cluster above)
    }
    else if (reduced)
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
        float2 w = TWS ? stw[k] : __ldg(&tw[k]);
        w.y = -w.y;  // w^{-k}
        const float tr = -(dr * w.y + di * w.x), ti = dr * w.x - di * w.y;
        b[k] = make_float2(er + tr, ei + ti);
        if (k2 != k) b[k2] = make_float2(er - tr, -(ei - ti));
      }
    }
    group_sync(g + 1);
    const float2* z = TWS ? fft_smem<true, true>(b, a, B, stw, t, g + 1) : fft_smem<true, false>(b, a, B, tw, t, g + 1);
    // keep the second half: time index i = 2j, 2j+1 with j >= B/2
#pragma unroll
    for (int q = 0; q < MAXRES; ++q)
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

  SSR_INV_STAMP(5)
  // groups 1.. hand their faded blocks to group 0 through their own smem region
  if (ngroups > 1)
  {
    if (g > 0)
    {
#pragma unroll
      for (int q = 0; q < MAXRES; ++q)
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
        for (int q = 0; q < MAXRES; ++q)
        {
          const int jj = t + q * kThreads;
          if (jj < half) { const float2 v = other[jj]; res[q].x += v.x; res[q].y += v.y; }
        }
      }
  }

  if (ga.consumed)
  {
    __syncthreads();          // thread 0's slot check precedes every store
    if (!s_slot_ok) return;   // timed out: the slot still holds an unread block -- store nothing, signal nothing
  }
  if (g == 0)
  {
    float m = 0.0f;
    float2* out = reinterpret_cast<float2*>(job.out + ga.out_offset);
#pragma unroll
    for (int q = 0; q < MAXRES; ++q)
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
  SSR_INV_STAMP(6)
#undef SSR_INV_STAMP
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
  if (ga.host_flag)
  {
    // all stores of this CTA (host memory, posted writes over PCIe) -> fence -> count
    // on a device-local counter; the CTA that comes last publishes the rank's flag
    // behind everybody's stores (fence cumulativity; PCIe keeps posted writes in order)
    __syncthreads();
    if (threadIdx.x == 0)
    {
      __threadfence_system();
      const unsigned prev = atomicAdd(ga.cta_counter, 1u);
      if (prev + 1 == ga.cta_total)
      {
        *ga.cta_counter = 0;   // every other CTA has counted in: ready for the next launch
        __threadfence_system();
        asm volatile("st.release.sys.global.u64 [%0], %1;" :: "l"(ga.host_flag), "l"(ga.flag_value) : "memory");
      }
    }
  }
}

template<int GROUPS>
__global__ void __launch_bounds__(kThreads * GROUPS, GROUPS == 1 ? 4 : 1)
inv_fft_kernel(const InvJob* __restrict__ jobs, const InvEntry* __restrict__ entries,
               const float2* __restrict__ partials, const float2* __restrict__ reduced, int B,
               const float2* __restrict__ tw,
               const float* __restrict__ fade_out, const float* __restrict__ fade_in,
               const GatherArgs ga, const int* __restrict__ kind_counts, const HeadCommit hc)
{
  if (inv_prologue(hc, kThreads * GROUPS)) return;
  inv_fft_body<GROUPS, false>(jobs, entries, partials, reduced, B, tw, fade_out, fade_in, ga, kind_counts, hc, 0);
}

// 4096 < B <= 8192 (a tail level of the non-uniformly partitioned renderer): one entry group, no staged twiddles
__global__ void __launch_bounds__(kThreads, 1)
inv_fft_large_kernel(const InvJob* __restrict__ jobs, const InvEntry* __restrict__ entries,
                     const float2* __restrict__ partials, const float2* __restrict__ reduced, int B,
                     const float2* __restrict__ tw,
                     const float* __restrict__ fade_out, const float* __restrict__ fade_in,
                     const GatherArgs ga, const int* __restrict__ kind_counts, const HeadCommit hc)
{
  if (inv_prologue(hc, kThreads)) return;
  inv_fft_body<1, false, kInvMaxResLarge, false>(jobs, entries, partials, reduced, B, tw, fade_out, fade_in, ga, kind_counts, hc, 0);
}

// Binaural / BRS behind mac_dyn_tma_kernel: grid = outputs x cluster size (launch attribute), three
// kThreads-thread groups = the three accumulator kinds; `partials` = [dyn_G][kind][ear][B]
__global__ void __launch_bounds__(kThreads * 3, 1)
inv_dyn_cluster_kernel(const InvJob* __restrict__ jobs, const InvEntry* __restrict__ entries,
                       const float2* __restrict__ partials, int dyn_G, int B, const float2* __restrict__ tw,
                       const float* __restrict__ fade_out, const float* __restrict__ fade_in,
                       const int* __restrict__ kind_counts, const HeadCommit hc, unsigned long long* dbg,
                       int early_trigger)
{
  // the next block's control upload (ctl_upload_kernel) does not depend on this kernel: let it in
  // early_trigger == 2 (block overlap): the launch carries one EXTRA CLUSTER whose leader commits the ring heads
  // first (the MAC, their last reader, is complete when this kernel starts); only then has the whole grid
  // triggered, and the next block's forward transform / control upload / MAC may run under this kernel
  if (early_trigger == 2)
  {
    unsigned crank, csize;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(crank));
    asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(csize));
    if (blockIdx.x / csize == gridDim.x / csize - 1)
    {
      if (crank == 0)
      {
        for (int i = threadIdx.x; i < hc.n; i += kThreads * 3)
        {
          int* head = hc.jobs[i].head;
          int h = *head + 1;
          if (h >= hc.jobs[i].parts) h = 0;
          *head = h;
        }
        __threadfence();
        __syncthreads();
      }
      grid_launch_dependents();
      return;
    }
    grid_launch_dependents();
  }
  else if (early_trigger) grid_launch_dependents();
  inv_fft_body<3, true>(jobs, entries, partials, nullptr, B, tw, fade_out, fade_in, GatherArgs{}, kind_counts, hc, dyn_G, dbg);
}

}  // namespace ssrk
