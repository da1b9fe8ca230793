// engine.hpp -- host side of the B200 partitioned-convolution engine.
//
// The classes here own HBM-resident spectra and build the small job tables the
// kernels in kernels.cuh consume.  They mirror the reference's object model
// (apf/apf/convolver.h: Filter / Input / Output / StaticOutput) but batch the
// arithmetic: one launch handles every (source, output, partition) triple of a
// block instead of one call per pair.
#pragma once

#include <cuda_runtime.h>

#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <mutex>
#include <cstdint>
#include <cstring>
#include <deque>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/ssr_b200.h"
#include "kernels.cuh"

namespace ssr_b200
{

struct Error : std::runtime_error
{
  Error(int code_, const std::string& what) : std::runtime_error(what), code(code_) {}
  int code;
};

#define SSR_CUDA(call) \
  do { cudaError_t err__ = (call); if (err__ != cudaSuccess) \
    throw ::ssr_b200::Error(SSR_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(err__)); \
  } while (0)

// "Layout epoch": bumped whenever a device buffer moves or a device table changes its
// length -- anything a captured CUDA graph has baked in (pointers, grid sizes).  Table
// CONTENT may change without invalidating a graph.  The counter is PER ENGINE: every
// entry point selects its engine with Engine::use_device(), which sets this thread's
// epoch slot, so one engine's allocations never invalidate another engine's graphs
// (engines beyond 63 share slots, which is only conservative).  Allocations made outside
// any engine scope count on slot 0, which every graph checks as well.
constexpr int kEpochSlots = 64;
inline std::atomic<uint64_t> g_layout_epochs[kEpochSlots];
inline std::atomic<unsigned> g_next_epoch_slot{0};
inline thread_local int t_epoch_slot = 0;
inline void bump_layout_epoch() { ++g_layout_epochs[t_epoch_slot]; }

// ------------------------------------------------------------ memory helpers
template<typename T>
struct DeviceArray
{
  T* ptr = nullptr;
  size_t count = 0;
  DeviceArray() {}
  explicit DeviceArray(size_t n) { resize(n); }
  DeviceArray(const DeviceArray&) = delete;
  DeviceArray& operator=(const DeviceArray&) = delete;
  ~DeviceArray() { release(); }
  void release() { if (ptr) { cudaFree(ptr); bump_layout_epoch(); } ptr = nullptr; count = 0; }
  void resize(size_t n)
  {
    release();
    if (n == 0) return;
    bump_layout_epoch();
    cudaError_t err = cudaMalloc(&ptr, n * sizeof(T));
    if (err != cudaSuccess)
      throw Error(err == cudaErrorMemoryAllocation ? SSR_ERR_NOMEM : SSR_ERR_CUDA,
          std::string("cudaMalloc: ") + cudaGetErrorString(err));
    count = n;
  }
  void ensure(size_t n) { if (n > count) resize(n + n / 2 + 16); }
};

template<typename T>
struct PinnedArray
{
  T* ptr = nullptr;
  size_t count = 0;
  PinnedArray() {}
  explicit PinnedArray(size_t n) { resize(n); }
  PinnedArray(const PinnedArray&) = delete;
  PinnedArray& operator=(const PinnedArray&) = delete;
  ~PinnedArray() { release(); }
  void release() { if (ptr) cudaFreeHost(ptr); ptr = nullptr; count = 0; }
  void resize(size_t n)
  {
    release();
    if (n == 0) return;
    SSR_CUDA(cudaMallocHost(&ptr, n * sizeof(T)));
    count = n;
  }
  void ensure(size_t n) { if (n > count) resize(n + n / 2 + 16); }
};

// A host-built table mirrored on the device (pinned staging + async copy).
template<typename T>
struct Table
{
  std::vector<T> host;
  PinnedArray<T> stage;
  DeviceArray<T> dev;
  size_t uploaded = 0;
  T* device() const { return dev.ptr; }
  cudaEvent_t staged = nullptr;   // recorded after the last staging -> device copy
  Table() {}
  Table(const Table&) = delete;
  Table& operator=(const Table&) = delete;
  ~Table() { if (staged) cudaEventDestroy(staged); }
  // Asynchronous on `s`.  The staging buffer and the device copy are reused, so
  // this first waits (on the host) until the previous upload has executed, and
  // (growing the device buffer) until everything queued on `s` is done: kernels
  // already queued may still read the old table.
  void upload(cudaStream_t s)
  {
    if (uploaded != host.size()) bump_layout_epoch();
    uploaded = host.size();
    if (host.empty()) return;
    if (staged) SSR_CUDA(cudaEventSynchronize(staged));
    else SSR_CUDA(cudaEventCreateWithFlags(&staged, cudaEventDisableTiming));
    if (host.size() > dev.count || host.size() > stage.count)
    {
      SSR_CUDA(cudaStreamSynchronize(s));
      stage.ensure(host.size());
      dev.ensure(host.size());
    }
    std::memcpy(stage.ptr, host.data(), host.size() * sizeof(T));
    SSR_CUDA(cudaMemcpyAsync(dev.ptr, stage.ptr, host.size() * sizeof(T),
          cudaMemcpyHostToDevice, s));
    SSR_CUDA(cudaEventRecord(staged, s));
  }
};

struct Engine;
struct FilterSet;
struct Input;

// ------------------------------------------------------------ MAC plan
// Rows of a group share their term sequence (and therefore every X load).
struct MacPlan
{
  int MT = 1;
  Table<ssrk::MacTermMeta> meta;
  Table<const float4*> hptr;   // [nterms][MT]
  // The term table (all groups back to back) is cut into `ncta` equal shares, one per
  // CTA; a share that straddles a group boundary becomes two (or more) SEGMENTS, each
  // with its own partial spectra.  work = segments, CTA c owns segments
  // [cta_first[c], cta_first[c + 1]) -- its longest one first.
  Table<int4> work;            // x = term_begin, y = term_end, z = partial index
  Table<int> cta_first;        // [ncta + 1]
  Table<int4> cta_head;        // [ncta][2]: {cta_first[c], cta_first[c + 1], 0, 0}, work[cta_first[c]] (bulk-copy MAC)
  int ncta = 0;
  int npartials = 0;           // partial spectra groups written per launch (= segments)
  struct Group { int term_begin, term_end, partial_first, nsplit; };
  std::vector<Group> groups;
  void clear()
  { meta.host.clear(); hptr.host.clear(); work.host.clear(); cta_first.host.clear(); cta_head.host.clear(); groups.clear(); npartials = 0; ncta = 0; }
  int begin_group() { groups.push_back({int(meta.host.size()), 0, 0, 0}); return int(groups.size()) - 1; }
  void add_term(int input, int p, int wslot, const float4* const* h)
  {
    meta.host.push_back({input, p, wslot, 0});
    for (int j = 0; j < MT; ++j) hptr.host.push_back(h[j]);
  }
  void end_group() { groups.back().term_end = int(meta.host.size()); }
  // `slots` = CTAs resident at once; `fixed_cost` = cost of one more CTA in units of one
  // term (pipeline fill, partials); `waves` > 0 forces waves x slots CTAs (experiments)
  // `weights` (one per CTA of the resulting plan, or null): relative speed of each CTA -- its share of
  // the term table is proportional to it (Renderer::calibrate)
  void make_work(int slots, int fixed_cost = 2, int waves = 0, const std::vector<double>* weights = nullptr);
  void upload(cudaStream_t s) { meta.upload(s); hptr.upload(s); work.upload(s); cta_first.upload(s); cta_head.upload(s); }
};

struct InvPlan
{
  Table<ssrk::InvJob> jobs;
  Table<ssrk::InvEntry> entries;
  void clear() { jobs.host.clear(); entries.host.clear(); }
  void upload(cudaStream_t s) { jobs.upload(s); entries.upload(s); }
};

// ------------------------------------------------------------ engine
struct StageEvents { cudaEvent_t ev[4]; uint64_t mac_bytes, mac_terms; int mac_launches, fft_launches, ifft_launches; };

struct Engine
{
  explicit Engine(const ssr_engine_config& cfg);
  ~Engine();

  int device, B, sample_rate, sm_count, mac_variant;
  static constexpr int kMaxBlock = 8192;
  bool large() const { return B > 4096; }   // transform buffers of 128 KiB: kernels without staged twiddles
  bool fft1024 = false;            // B == 1024: register/shuffle FFT kernels (fft1024.cuh)
  bool fft1024_always = false;
  bool pdl = true;                 // programmatic dependent launch of the MAC / reduce / inverse kernels
  void launch_mac_fn(void* fn, dim3 grid, void** args, int threads = ssrk::kThreads, size_t smem = 0);
  // offline: kTimeBatch blocks per launch (mac_tma_batch_kernel); partials [CTA][t][row][B]
  void launch_mac_batch(const MacPlan& plan, int wtab_offset);
  bool mac_tma() const;
  bool level1_fused = true;        // Level 1: one launch per convolve() for short term lists
  void launch_level1_fused(const ssrk::FusedParams& p);
  static constexpr int kFft1024MinJobs = 1024;
  cudaStream_t stream;
  bool own_stream;

  DeviceArray<float2> tw;          // exp(-2 pi i k / 2B), k in [0, 2B)
  DeviceArray<float> fade_out, fade_in;
  DeviceArray<float2> zero_part;   // B float2 of zeros ("empty partition", convolver.h:415)

  // input registry (device tables indexed by input id)
  static constexpr int kMaxInputs = 16384;
  DeviceArray<const float4*> d_xring;
  DeviceArray<int> d_xparts, d_heads;
  std::vector<Input*> inputs;
  int register_input(Input* in);
  void unregister_input(int id);

  // bumped whenever a zero flag of any filter set of this engine may have changed; plans
  // that snapshot "partition present / absent" (Output::convolve) key on it
  uint64_t flags_epoch = 1;
  Table<float> wtab;               // weight slots
  DeviceArray<float4> partials;
  DeviceArray<float4> reduced;     // pre-reduced accumulators (see reduce_partials_kernel)

  // launches (all asynchronous on `stream`)
  // overlap: the launch takes part in the programmatic-dependent-launch chain of a block
  // loop and leaves the ring heads to be committed by the block's inverse kernel
  // (ssrk::HeadCommit), so that the MAC launched behind it may run concurrently
  // late_wait: see fwd_fft_overlap_kernel (block overlap; only with `overlap`)
  // dyn_variant: the function variant that runs next to the generated-terms bulk-copy MAC, without a ring update
  // (block overlap of the Binaural / BRS chain: the update moved into ctl_upload_kernel)
  void launch_fwd(const ssrk::FwdJob* d_jobs, int njobs, const ssrk::FwdRingUpdate* ring_update = nullptr,
                  bool overlap = false, const float* in_base = nullptr, bool late_wait = false,
                  bool dyn_variant = false);
  void launch_fwd_batch(const ssrk::FwdBatch& batch, int nparts, int nchannels);
  void launch_mac(const MacPlan& plan, int wtab_offset, int partial_offset, const float* weights = nullptr,
                  bool fwd_overlap = false);
  void launch_inv(const InvPlan& plan, const ssrk::GatherArgs& ga = ssrk::GatherArgs{},
                  const ssrk::HeadCommit& hc = ssrk::HeadCommit{});
  bool fwd_overlap = true;         // forward transform under the MAC (SSR_B200_FWD_OVERLAP=0: off)
  // Block overlap (SSR_B200_BLOCK_OVERLAP): in a loop of device-resident blocks the NEXT block's forward transform
  // and the p != 0 terms of its MAC run under THIS block's inverse transform (and, on the root of a sharded
  // renderer, under the gather's wait / release helpers).  The inverse kernel commits the ring heads before it
  // lets its dependents in (HeadCommit::early_trigger == 2), the forward kernel passes the trigger on at once and
  // waits at its end (late_wait); the MAC's p == 0 terms and its first partial-spectra store still wait for the
  // forward kernel's completion, i.e. for the previous block's inverse transform.  Generic plans on the
  // bulk-copy MAC; a synchronous host call (one block at a time) is unaffected.
  bool block_overlap = true;
  int mac_waves = 0;               // SSR_B200_MAC_WAVES: CTAs of a MAC launch = waves x resident CTAs
  DeviceArray<unsigned long long> mac_trace;   // per-CTA timestamps / busy time of the bulk-copy MAC launches
  bool early_trigger = false;                  // SSR_B200_EARLY_TRIGGER=1: the bulk-copy MACs let their dependent kernel launch at once
                                               //  (measured: the block gets 5 us SLOWER on cfg4, profiles/r2_ab; off)
  bool inv_trigger = false;                    // SSR_B200_INV_TRIGGER=1: see ssrk::HeadCommit::early_trigger
  int trace_mode = 0;                          // SSR_B200_TRACE_ISSUE=1: trace[3] = first bulk copy (see MacParams)
  bool mac_calibrate = false;                  // SSR_B200_MAC_CALIBRATE=1: re-cut the shares by measured CTA rates
  int prereduce_min = 16;          // partials per accumulator from which a separate reduction launch pays
  // dynamic renderers: terms generated on the device (ssrk::DynTables), static launch shapes
  void launch_mac_dyn(const ssrk::DynTables& dyn, int G, const float* weights, bool fwd_overlap = false);
  void launch_inv_dyn(const InvPlan& plan, const ssrk::DynCtl* ctl, int Pmax, int G,
                      const ssrk::HeadCommit& hc = ssrk::HeadCommit{}, bool early_trigger = false);
  // ring_update != null: late-trigger mode (behind the forward transform, see ctl_upload_kernel)
  void launch_ctl_upload(const void* src_host, void* dst, int n16, unsigned long long* counter,
                         unsigned long long* ack, const ssrk::FwdRingUpdate* ring_update = nullptr, int ring_rows = 0);
  void launch_dyn_ring_update(const ssrk::DynCtl* ctl, const ssrk::DynRingEntry* upd,
                              ssrk::DynRingEntry* ring, int rows, int R);
  int dyn_mac_slots() const;
  bool dyn_tma_enabled = true;     // SSR_B200_DYN_TMA=0: the register-staged generated-terms MAC + dyn_reduce_kernel
  bool dyn_tma() const;            // Binaural / BRS on mac_dyn_tma_kernel + inv_dyn_cluster_kernel
  int dyn_cluster = 8;             // CTAs per cluster of inv_dyn_cluster_kernel
  int inv_groups(int max_entries) const;
  void launch_inv_kernel(int groups, int njobs, const InvPlan& plan, const float2* red,
                         const ssrk::GatherArgs& ga, const int* kind_counts,
                         const ssrk::HeadCommit& hc = ssrk::HeadCommit{});       // CTAs resident at once for the generated-terms MAC kernel
  DeviceArray<unsigned long long> d_dyn_stats;  // [2] H / X partitions read by generated-terms launches
  uint64_t dyn_h_seen = 0, dyn_x_seen = 0;
  void upload_weights(const float* w, size_t n);
  void ensure_partials(size_t npartial_spectra);
  int mac_slots(int MT) const;     // CTAs resident at once for the MAC kernel of MT-row plans
  static constexpr int kSplitCostTma = 8;
  int split_cost(int MT) const;    // see MacPlan::make_work

  // stats
  bool timing = false;
  int timing_every = 1;            // stage events on every timing_every-th block step
  uint64_t timing_counter = 0;
  void count_launch(int n = 1) { stats.launches_total += uint64_t(n); }
  ssr_engine_stats stats{};
  std::vector<StageEvents> pending;
  std::vector<StageEvents> free_events;
  StageEvents* cur = nullptr;
  void stage_begin();              // call before the first kernel of a block step
  void stage_mark(int i);          // 1 after fft, 2 after mac, 3 after ifft
  void stage_end();
  void fold_stats();
  void sync() { SSR_CUDA(cudaStreamSynchronize(stream)); }
  const int epoch_slot = 1 + int(g_next_epoch_slot++ % unsigned(kEpochSlots - 1));
  uint64_t graph_epoch() const { return g_layout_epochs[epoch_slot].load() + g_layout_epochs[0].load(); }
  void use_device() { SSR_CUDA(cudaSetDevice(device)); t_epoch_slot = epoch_slot; }

  // Is [p, p + bytes) page-locked host memory (cudaMallocHost / cudaHostRegister)?
  // Answers are cached by address: a stale "pinned" answer only costs speed, since
  // cudaMemcpyAsync stays correct on pageable memory.
  struct PinnedRange { const void* p; size_t bytes; bool pinned; };
  std::vector<PinnedRange> pinned_cache;
  bool is_pinned(const void* p, size_t bytes);
};

// ------------------------------------------------------------ filter sets
struct FilterSet
{
  FilterSet(Engine* e, int channels, int partitions);
  Engine* e;
  int channels, partitions;
  DeviceArray<float2> data;          // [channels][partitions][B]
  std::vector<int> valid;            // per channel: partitions [0, valid) carry data
  std::vector<uint8_t> flags;        // per (channel, partition): 1 = non-zero-flagged
  DeviceArray<unsigned char> d_flags;  // device mirror of `flags` (generated MAC terms read it)
  bool flags_dirty = true;
  void touch_flags();                  // a zero flag may have changed: device mirror and cached plans are stale
  const unsigned char* device_flags(); // uploads when stale (stream-ordered)
  std::vector<uint8_t> all_set;        // per channel: no zero-flagged partition
  // device flags of one channel, or null when every partition carries data
  const unsigned char* device_flags_row(int ch);
  const float2* part(int ch, int p) const { return data.ptr + (size_t(ch) * partitions + p) * e->B; }
  float2* part(int ch, int p) { return data.ptr + (size_t(ch) * partitions + p) * e->B; }
  // null for zero-flagged partitions (fft_node::zero, convolver.h:91-96)
  const float2* part_or_null(int ch, int p) const
  { return (p < partitions && flags[size_t(ch) * partitions + p]) ? part(ch, p) : nullptr; }
  void load_time(int first_channel, int nch, const float* data, long frames,
                 long frame_stride, long channel_stride);
  void set_partition_time(int ch, int p, const float* taps, int n);
  // tap_offset: the set holds taps [tap_offset, tap_offset + taps) of the generated responses (a level of
  // the non-uniformly partitioned renderer); `envelope` then points at that slice's first entry
  void generate(int first_channel, int nch, long taps, uint64_t seed, uint64_t id_offset,
                const float* envelope, float scale, uint64_t tap_offset = 0);
  void lerp_from(int dch, const FilterSet& a, int ach, const FilterSet& b, int bch, float t);
  // host half of lerp_from: zero flags + one job per partition appended to `jobs`; the caller
  // launches ssrk::lerp_kernel over them (the renderers batch a block's interpolations
  // into one stream-ordered launch, no allocation, no synchronisation)
  void lerp_plan(int dch, const FilterSet& a, int ach, const FilterSet& b, int bch, float t,
                 std::vector<ssrk::LerpJob>& jobs);
  void download(int ch, int p, float* sorted_out, int* zero);
  void upload(int ch, int p, const float* sorted_in, int zero);
};

// pointer table + staggered switch queues of one dynamic Output
// (apf/apf/convolver.h:618-699); pointers are device addresses, null = empty
struct FilterPtrTable
{
  explicit FilterPtrTable(int partitions);
  int P;

  // The reference keeps P-1 shifting queues of lengths 1..P-1 (O(P^2) slots and
  // pointer moves per rotate, convolver.h:623-624,690-698).  All queues are fed
  // together by set_filter(), so the same behaviour follows from a short list of
  // switch EVENTS: partition p of a filter installed after T_set rotations is
  // used once T >= T_set + p, i.e.
  //     ptr[p](T) = part p of the newest event with T_set + p <= T   (else base[p]).
  // rotate_queues() is then O(1) and the table is materialised on demand in
  // O(P + events).
  struct Event
  {
    uint64_t t_set;
    const float2* base; size_t stride; const uint8_t* flags; int nparts;   // a FilterSet channel ...
    std::shared_ptr<std::vector<const float2*>> list;                       // ... or an explicit list
    const float2* part(int p) const
    {
      if (list) return p < int(list->size()) ? (*list)[p] : nullptr;
      return (p < nparts && flags[p]) ? base + size_t(p) * stride : nullptr;  // beyond -> empty
    }
  };
  uint64_t T = 0;                                  // rotate_queues() calls so far
  // table before / below all pending events: a filter (evaluated LIVE, so that a zero
  // flag flipped in place after binding is seen -- the reference reads fft_node::zero in
  // every _multiply_spectra, convolver.h:556-561), or nothing (all partitions empty)
  Event base{0, nullptr, 0, nullptr, 0, nullptr};
  std::vector<Event> events;                       // ascending t_set, at most one per T

  void set_static(const FilterSet& fs, int ch);    // StaticOutput::_set_filter :729-739
  void set_filter(const FilterSet& fs, int ch);    // :642-658
  // the same on a plain list of partition pointers (null = zero-flagged,
  // entries beyond the list = empty partition); pure host logic
  void set_static_ptrs(const std::vector<const float2*>& parts);
  void set_filter_ptrs(const std::vector<const float2*>& parts);
  uint64_t version = 0;                            // bumped whenever the table may have changed
  bool queues_empty() const;                       // :666-677
  void rotate_queues();                            // :683-699
  // _filter_ptrs as of now (null = empty / zero-flagged partition)
  const std::vector<const float2*>& current();
  bool references(const float2* lo, const float2* hi) const;

  private:
    std::vector<const float2*> _ptrs;
    uint64_t _ptrs_version = ~uint64_t(0);
    void push_event(Event ev);
};

// ------------------------------------------------------------ Level 1 objects
struct Input
{
  Input(Engine* e, int partitions);
  ~Input();
  Engine* e;
  int id, P;
  // delay-line slots: P partitions + (kTimeBatch - 1) spare ones, so that the
  // offline time-batched MAC can read kTimeBatch consecutive blocks' spectra
  static constexpr int kTimeBatch = 4;
  int slots;
  DeviceArray<float2> ring;          // [slots][B]
  DeviceArray<float> prev, cur;      // B floats each
  PinnedArray<float> stage;
  Table<ssrk::FwdJob> job;
  // add_block() only stages the block; the forward transform runs inside the first
  // Output::convolve() of the block (fused launch) or when flush() is called
  bool fwd_pending = false;
  bool stage_busy = false;           // an asynchronous copy may still read the staging buffer
  void flush();                      // run a pending forward transform now (asynchronous)
  ssrk::FwdJob make_job(const float* d_cur);
  void add_block(const float* x);
  void download(int p, float* sorted_out);
};

struct Output
{
  Output(Input* in, ssr_output_kind kind);
  Input* in;
  Engine* e;
  ssr_output_kind kind;
  FilterPtrTable table;
  MacPlan mac;
  InvPlan inv;
  uint64_t planned_version = ~uint64_t(0);   // table.version the cached plan was built from
  uint64_t planned_flags_epoch = 0;          // Engine::flags_epoch it was built under
  int planned_terms = 0;
  bool planned_zero = false;
  DeviceArray<float> d_res;
  PinnedArray<float> h_res;
  PinnedArray<int> h_flag;           // completion flag of the fused launch (mapped host memory)
  int fused_seq = 0;
  const float* convolve(float weight);
  const float* convolve_fused(float weight);
};

// ------------------------------------------------------------ multi-GPU output gather
// The path's only exchange step: every rank's M/G x B output blocks go
// to the host-facing rank.  Instead of a collective after the inverse kernel, the
// inverse kernel of a peer rank stores its blocks straight into the root's gather
// buffer through NVLink peer memory and counts its CTAs in on a system-scope
// counter (ssrk::GatherArgs); the root only runs a one-warp wait kernel.
//
// One device allocation on the root (exported with cudaIpcGetMemHandle, or shared
// by pointer between engines of one process):
//   [control block 4 KiB: consumed | arrived[rank] (one 128-byte line each)]
//   [slot 0: M_total x B floats][slot 1: M_total x B floats]
struct Gather
{
  Gather(Engine* e, int world, int rank, int root, const int* outputs_per_rank);
  ~Gather();
  Engine* e;
  int world, rank, root;
  std::vector<int> counts, firsts;   // outputs per rank, first output of each rank
  int M_total = 0;
  bool is_root() const { return rank == root; }

  static constexpr size_t kControlBytes = 4096;
  static constexpr int kMaxWorld = 24;
  unsigned char* base = nullptr;     // root: own allocation; peer: mapped pointer
  bool owns = false, ipc_mapped = false;
  // local sticky time-out flag in MAPPED PINNED host memory: written by the waiting
  // kernels, read by the host after a block's completion event at no cost
  PinnedArray<int> h_error;
  int* d_error = nullptr;            // the device's view of h_error
  bool timed_out() const { return h_error.ptr && *reinterpret_cast<volatile int*>(h_error.ptr) != 0; }
  // ranks driven by one host thread on ONE device (single-GPU tests): see ssr_b200.h
  bool shares_root_device = false;
  unsigned long long timeout_ns = 0;  // 0: ssrk::kGatherTimeoutNs (20 s)
  DeviceArray<int> d_ctas;           // root: CTAs per rank (0 for the root)
  uint64_t epoch = 0;                // blocks rendered into the buffer so far
  uint64_t waited = 0, released = 0; // root bookkeeping

  size_t slot_floats() const { return size_t(M_total) * e->B; }
  size_t total_bytes() const { return kControlBytes + 2 * slot_floats() * sizeof(float); }
  unsigned long long* consumed_ptr() const { return reinterpret_cast<unsigned long long*>(base); }
  unsigned long long* arrived_ptr(int g) const
  { return reinterpret_cast<unsigned long long*>(base) + size_t(1 + g) * ssrk::kGatherCounterStride; }
  float* slot_ptr(int slot) const
  { return reinterpret_cast<float*>(base + kControlBytes) + size_t(slot) * slot_floats(); }

  void export_handle(void* handle64) const;            // root
  void import_handle(const void* handle64);            // peer, other process (CUDA IPC)
  void import_pointer(void* root_base, int root_device);  // peer, same process
  bool ready() const { return base != nullptr; }
  // kernel arguments for the block about to be rendered (epoch), then ++epoch
  ssrk::GatherArgs next_block();
  // root: stream-ordered wait for block (epoch - 1) of every rank; returns its slot
  const float* wait();
  // root: after the reads of that slot were queued (on `s`, default the engine's stream)
  void release(cudaStream_t s = nullptr);
  int status();                                        // 0 ok, 1 a wait timed out (synchronises)
};

// ------------------------------------------------------------ host-direct output delivery
// Second realisation of the path's exchange step, for the HOST-buffer call
// (pointer_policy::audio_callback, apf/apf/pointer_policy.h:112-122, where the
// reference's Output items write the caller's host channels directly): one POSIX
// shared-memory segment, page-locked and device-mapped in every rank
//   [control 4 KiB: consumed | flag[rank] (one 128-byte line each)]
//   [slot 0: M_total x B floats][slot 1: M_total x B floats]
// Rank g's inverse kernel stores its M/G x B blocks into slot (t & 1) over ITS OWN
// PCIe link (posted writes that overlap the kernel) and its last CTA then writes
// flag[g] = t + 1.  The root's host thread returns from the block when every flag
// is in; no GPU ever waits for another one, no 2 MB D2H copy funnels through the
// root's link, and the caller reads the output blocks in place (hostgather.cu).
struct HostGather
{
  HostGather(Engine* e, int world, int rank, int root, const int* outputs_per_rank, const char* shm_name);
  ~HostGather();
  HostGather(const HostGather&) = delete;
  HostGather& operator=(const HostGather&) = delete;
  Engine* e;
  int world, rank, root;
  std::string name;
  std::vector<int> counts, firsts;
  int M_total = 0;
  bool is_root() const { return rank == root; }

  static constexpr size_t kControlBytes = 4096;
  static constexpr int kMaxWorld = 24;
  unsigned char* base = nullptr;     // host mapping of the segment
  unsigned char* d_base = nullptr;   // the device's view of it
  size_t bytes = 0;
  bool created = false, registered = false, failed = false;
  double timeout_ms = 20000.0;
  DeviceArray<unsigned> d_counter;   // CTAs of the running inverse launch that have stored their block
  uint64_t epoch = 0;                // blocks rendered into the segment so far

  size_t slot_floats() const { return size_t(M_total) * e->B; }
  float* slot_ptr(int slot) const { return reinterpret_cast<float*>(base + kControlBytes) + size_t(slot) * slot_floats(); }
  float* d_slot_ptr(int slot) const { return reinterpret_cast<float*>(d_base + kControlBytes) + size_t(slot) * slot_floats(); }
  unsigned long long* consumed_ptr() const { return reinterpret_cast<unsigned long long*>(base); }
  volatile unsigned long long* flag_ptr(int g) const
  { return reinterpret_cast<volatile unsigned long long*>(base) + size_t(1 + g) * ssrk::kGatherCounterStride; }

  ssrk::GatherArgs next_block(int ctas);
  void publish_consumed(uint64_t t);
  void wait_slot_free(uint64_t t);
  void wait_all(uint64_t t);
};

// ------------------------------------------------------------ Level 2
template<typename T>
struct BlockParameter  // apf::BlockParameter (apf/apf/misc.h:94-195)
{
  T current, old_value;
  explicit BlockParameter(T v = T()) : current(v), old_value(v) {}
  void assign(T v) { old_value = current; current = v; }
  const T& get() const { return current; }
  const T& old() const { return old_value; }
  bool changed() const { return current != old_value; }
  bool both_eq(T v) const { return current == v && old_value == v; }
};

enum Mode { NOTHING = 0, CONSTANT = 1, CHANGE = 2, FADE_IN = 3, FADE_OUT = 4 };

// SSR_B200_PROFILE_HOST=1: wall time of the host-side sections of a block,
// printed when the renderer is destroyed (development aid)
struct HostProfile
{
  bool on = std::getenv("SSR_B200_PROFILE_HOST") != nullptr;
  double acc[8] = {0}; uint64_t n = 0;
  std::chrono::steady_clock::time_point t;
  void start() { if (on) t = std::chrono::steady_clock::now(); }
  void lap(int i)
  {
    if (!on) return;
    auto now = std::chrono::steady_clock::now();
    acc[i] += std::chrono::duration<double, std::micro>(now - t).count();
    t = now;
  }
  void report(const char* what)
  {
    if (!on || !n) return;
    std::printf("[ssr_b200 host profile] %s: %llu blocks; us/block:", what, (unsigned long long)n);
    for (int i = 0; i < 8; ++i) if (acc[i] > 0) std::printf(" s%d=%.1f", i, acc[i] / n);
    std::printf("\n");
  }
};


// Control hand-off from non-realtime threads to the audio thread = apf::CommandQueue +
// apf::SharedData (apf/apf/commandqueue.h:141-153,185-210, apf/apf/shareddata.h:35-83):
// a setter called from a control thread (GUI, network, tracker; src/rendersubscriber.h:
// 164-169) does not touch the renderer's state, it queues a SetCommand; the audio thread
// executes everything queued, in order, at the START of its next block
// (apf/apf/mimoprocessor.h:374).  Single consumer (the thread inside process()), any
// number of producers (serialised by a mutex, like the reference's scoped lock).
struct ControlQueue
{
  enum Kind { SRC_GAIN, SRC_MUTE, SRC_POSITION, SRC_MODEL, REF_POSITION, REF_ORIENTATION, OFF_POSITION,
              OFF_ORIENTATION, MASTER_VOLUME, PROCESSING, AMP_REF_DISTANCE };
  struct Cmd { int kind; int id; float a, b; };
  static constexpr uint32_t kCapacity = 4096;     // (the reference's "fifo_size" defaults to 1024)
  Cmd ring[kCapacity];
  std::atomic<uint32_t> head{0}, tail{0};          // consumer / producer positions (free-running)
  std::mutex producers;
  // false: still full after ~0.2 s (nobody is rendering blocks)
  bool push(const Cmd& c);
  template<typename F> void drain(F apply)
  {
    uint32_t h = head.load(std::memory_order_relaxed);
    const uint32_t t = tail.load(std::memory_order_acquire);
    for (; h != t; ++h) apply(ring[h % kCapacity]);
    head.store(h, std::memory_order_release);
  }
};

struct Renderer
{
  Renderer(Engine* e, const ssr_renderer_config& cfg);
  ~Renderer();

  struct Source
  {
    int id;
    std::shared_ptr<Input> input;   // (shared by the sub-renderers of a tail level of the non-uniform renderer)
    float gain = 1.0f; bool mute = false; float px = 0, py = 0; bool plane = false;
    BlockParameter<float> weighting_factor{0.0f};
    // generic
    std::unique_ptr<FilterSet> irs;
    BlockParameter<float> w{0.0f};
    bool primed = false;       // tail_level: the first block's "old" weight has been set
    // brs / binaural
    int angles = 0;
    std::unique_ptr<FilterSet> brtf;
    BlockParameter<float> bw{-1.0f};
    BlockParameter<long> index{-1};
    BlockParameter<float> interp{-1.0f};
    // Staggered filter switch of the two ears' Outputs (convolver.h:618-699) as
    // "current filter by block index": latest = the filter set_filter() installed
    // last, hist[ear][b mod (P + 1)] = latest as of block b (mirror of the device ring)
    ssrk::DynRingEntry latest[2] = {{nullptr, nullptr, 0, 0}, {nullptr, nullptr, 0, 0}};
    std::vector<ssrk::DynRingEntry> hist[2];
    bool has_switched = false;
    long long last_switch = 0;                        // block index of the last set_filter()
    std::vector<std::vector<std::unique_ptr<FilterSet>>> temp;  // per ear: pool of temporary_hrtf
    Mode mode = NOTHING;
  };

  Engine* e;
  ssr_renderer_config cfg;
  bool skip_forward = false;   // the delay lines (Source::input) belong to another renderer of the same engine, which
                               //  has already transformed this block (non-uniform renderer: sub-renderers 1.. of a level)
  bool tail_level = false;     // a tail level of the non-uniformly partitioned renderer (nonuniform.cu): new
                               //  sources start at their weight instead of fading in from 0
  int M_total, m_first, m_local;
  std::vector<std::unique_ptr<Source>> sources;
  int next_id = 1;

  // RendererBase::State (src/rendererbase.h:84-103)
  float ref_x = 0, ref_y = 0, ref_az = 0, off_x = 0, off_y = 0, off_az = 0;
  float master_volume = 1.0f, amplitude_reference_distance = 3.0f, mvc = 1.0f;
  bool processing = true;

  // binaural
  std::unique_ptr<FilterSet> hrtfs, neutral;
  int angles = 0, partitions = 0;

  // per-block plans
  MacPlan static_mac;           // generic: built when the source list changes
  bool static_dirty = true;
  // Share calibration of the bulk-copy MAC (experiment, OFF by default: SSR_B200_MAC_CALIBRATE=1).
  // With one equal share per SM the CTAs of one launch finish up to 7 % apart (scripts/mac_trace.py:
  // 420 .. 475 us on the per-GPU shard of the 8-GPU run; the same CTAs are slow launch after launch).
  // The kernel accumulates each CTA's busy time; after a few steady blocks the shares are re-cut in
  // proportion to the measured rates, twice.  Measured: the spread barely shrinks (a CTA's rate depends
  // on what its neighbours do, not on its share alone) and the block gets 1 - 2 % faster at best, because
  // the CTAs still running keep HBM saturated during the tail anyway (profiles/r2_mac_balance.md).  Not
  // worth results that differ from run to run in the last bits, hence off.
  struct Calibration
  {
    int state = 0;                 // 0 counting steady blocks, 1 measuring, 2 read-back queued, 3 done
    int steady = 0, measured = 0, rounds = 0;
    std::vector<double> weights;   // per CTA; empty = equal shares
    PinnedArray<unsigned long long> host;
    cudaEvent_t ev = nullptr;
    void reset() { state = 0; steady = 0; measured = 0; rounds = 0; weights.clear(); }
  } calib;
  void calibrate(int kind_mask);
  InvPlan inv_plans[2];         // inverse tables per output buffer (pipelined submit() alternates)
  int inv_keys[2] = {-1, -1};
  int cur_slot = 0;
  void invalidate_inv();
  // brs / binaural: terms are generated on the device (ssrk::DynTables)
  long long dyn_b = 0;                   // block counter
  int dyn_Pmax = 1, dyn_R = 2, dyn_G = 0;
  DeviceArray<ssrk::DynRingEntry> dyn_ring;   // [sources][2 ears][R]
  // this block's control data: staged in page-locked memory, brought over by ctl_upload_kernel (or, with
  // SSR_B200_CTL_UPLOAD=copy, a stream-ordered copy); staging slot and device buffer alternate by block
  DeviceArray<unsigned char> dyn_dev[2];
  PinnedArray<unsigned char> dyn_stage[2];
  cudaEvent_t dyn_copied[2] = {nullptr, nullptr};
  int dyn_slot = 0;
  bool ctl_by_kernel = true;
  bool dyn_block_overlap = true;   // Binaural / BRS: block overlap with the chain forward -> control upload -> MAC
                                   //  (SSR_B200_DYN_BLOCK_OVERLAP=0: the chain control upload -> forward -> MAC)
  DeviceArray<unsigned long long> dyn_upload_count;      // [2] uploads done per staging slot (device)
  PinnedArray<unsigned long long> dyn_ack;               // [2] the same, as the host sees it
  unsigned long long dyn_uploads_issued[2] = {0, 0};
  ssrk::DynTables dyn_tables{};
  const float* dyn_wtab_dev = nullptr;
  const ssrk::DynRingEntry* dyn_upd_dev = nullptr;
  // forward-FFT job tables, one per distinct device input pointer (a caller
  // cycling through a few input buffers never re-uploads in steady state)
  std::vector<std::pair<const float*, std::unique_ptr<Table<ssrk::FwdJob>>>> fwd_cache;
  Table<ssrk::FwdJob>* fwd = nullptr;
  std::vector<float> wtab_host;
  Table<ssrk::LerpJob> lerp_tab;          // Binaural: this block's Dirac interpolations (one launch)
  DeviceArray<float> d_in[2], d_out[2];   // [1] only used by the pipelined submit()
  PinnedArray<float> h_in, h_out;

  HostProfile prof;                      // per renderer (development aid, off by default)
  ControlQueue control;
  void apply_controls();                 // audio thread, start of every block
  uint64_t blocks_started = 0;
  bool controls_applied = false;         // process_batch() already ran apply_controls() for the next block
  Source* find(int id);
  Source* find_or_null(int id);
  int add_source(std::unique_ptr<FilterSet> irs, int channels);
  void rem_source(int id);
  void base_weight(Source& s);
  // One block = prepare (host logic, table / weight uploads) + launch (kernel
  // launches only, capturable into a CUDA graph).
  struct StepDesc
  {
    const float* d_input = nullptr; float* d_output = nullptr;
    int N = 0;
    const float* in_base = nullptr;    // the caller's page-locked input array (device view), read in place
    int kind_mask = 0;                 // generic: accumulator kinds active (const / old / new)
    uint64_t hparts[3] = {0, 0, 0}, xparts[3] = {0, 0, 0};
    int nrows = 0;
    bool graphable = false;
    // Binaural / BRS: control upload done by the block's first kernel (ctl_upload_kernel)
    const void* ctl_src = nullptr; void* ctl_dst = nullptr; int ctl_n16 = 0; int ctl_slot = 0;
  };
  StepDesc prepare(const float* d_input, float* d_output);
  void launch(const StepDesc& d);
  void step(const float* d_input, float* d_output);   // async on e->stream
  StepDesc prepare_generic(const float* d_input, float* d_output);
  StepDesc prepare_dynamic(const float* d_input, float* d_output);
  StepDesc prepare_bank(const float* d_input, float* d_output);
  // CUDA graphs of the steady-state kernel sequence of a host-buffer step
  struct GraphEntry { cudaGraphExec_t exec = nullptr; uint64_t epoch = 0; int seen = 0; int kernels = 0; };
  GraphEntry graphs[2][8];             // [device-buffer slot][kind mask]
  bool use_graphs = true;
  bool zero_copy = true;               // process(): inverse kernel stores into the caller's pinned array
  bool zero_copy_in = true;            // process(): forward kernel reads the caller's pinned input array
  uint64_t graph_launches = 0;
  void drop_graphs();
  // level meters
  bool metering = false;
  DeviceArray<float> d_levels;   // [sources | local outputs]
  float* level_ptr(int index);
  void set_metering(bool on);
  void get_levels(float* pre_fader, float* levels, float* outputs);
  void build_fwd(const float* d_input);
  void process(int n, const float* const* in, float* const* out);
  // Offline: `nblocks` consecutive blocks (in[t * sources + i], out[t * outputs + m]);
  // runs of Input::kTimeBatch blocks with constant controls go through the
  // time-batched MAC, everything else through process()
  void process_batch(int n, int nblocks, const float* const* in, float* const* out);
  bool batch_eligible();
  void run_batch(const float* const* in, float* const* out);
  InvPlan inv_batch;
  int inv_batch_key = -1;
  DeviceArray<float> d_in_batch, d_out_batch;
  PinnedArray<float> h_in_batch, h_out_batch;
  uint64_t batched_blocks = 0;
  // multi-GPU: output blocks go to a Gather (not owned) instead of d_output
  Gather* gather = nullptr;
  void bind_gather(Gather* g);
  // ... or straight to the shared host segment of a HostGather (not owned)
  HostGather* hostgather = nullptr;
  void bind_hostgather(HostGather* g);
  // output blocks the host call delivers on this rank
  int host_outputs() const
  {
    if (gather) return gather->is_root() ? M_total : 0;
    if (hostgather) return hostgather->is_root() ? M_total : 0;
    return m_local;
  }
  // pipelined
  int outstanding = 0;
  uint64_t submitted = 0;
  cudaEvent_t done[2] = {nullptr, nullptr};
  bool slot_direct[2] = {false, false};  // the slot's D2H went straight into the caller's buffer
  uint64_t slot_epoch[2] = {0, 0};       // host-direct gather: block index the slot's call rendered
  void submit(int n, const float* const* in);
  void submit_block(int n, const float* const* in, float* const* direct_out, bool pipelined);
  // copy streams and events of the pipelined submit(): H2D | kernels | D2H
  cudaStream_t s_in = nullptr, s_out = nullptr;
  cudaEvent_t ev_in[2] = {nullptr, nullptr}, ev_k[2] = {nullptr, nullptr};
  bool slot_used[2] = {false, false};    // a pipelined block has used the slot's buffers before
  void wait(float* const* out);
};

}  // namespace ssr_b200
