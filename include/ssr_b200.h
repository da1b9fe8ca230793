/* ssr_b200.h -- C ABI of the B200-native partitioned-convolution engine.
 *
 * Drop-in boundary for ONE path of the SoundScape Renderer: the apf::conv
 * uniformly-partitioned overlap-save convolution engine and the per-block step
 * of the Binaural / BRS / Generic renderers that drive it.  Citations are
 * file:line in the reference tree (SoundScapeRenderer/ssr-pre-0.4.0).
 *
 * Conventions
 *  - plain C, opaque handles, plain pointers and sizes; no C++ / torch types.
 *  - every function returns SSR_OK (0) or a negative ssr_status; nothing throws
 *    across this boundary.  ssr_last_error() returns the message of the last
 *    failure on the calling thread.  The C++ facade
 *    (ssr-pre-0.4.0_b200/cpp/apf/convolver.h) rethrows the reference's exception
 *    types (std::logic_error, apf/apf/convolver.h:163-166).
 *  - there is NO CPU fallback: every compute entry point runs CUDA kernels for
 *    sm_100a and fails with SSR_ERR_CUDA when no device is usable.
 *  - "block" = audio block size B (samples); a partition spectrum is B complex
 *    bins = 2B floats (8 KiB at B = 1024), resident in HBM.
 *  - host result pointers are owned by the engine (pinned memory) and stay
 *    valid until the next call that produces a result for the same object.
 *  - threading: one engine is used by one host thread at a time (the reference
 *    gives the same guarantee per convolver object, apf/apf/convolver.h:53 and
 *    the list barrier apf/apf/mimoprocessor.h:476-482) -- EXCEPT the control
 *    setters ssr_source_set_* / ssr_renderer_set_*, which any thread may call at
 *    any time (command queue, see "Control inputs").
 */
#ifndef SSR_B200_H
#define SSR_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SSR_B200_ABI_VERSION 1

typedef enum ssr_status
{
  SSR_OK = 0,
  SSR_ERR_INVALID = -1,      /* bad argument / precondition (reference: assert) */
  SSR_ERR_BLOCK_SIZE = -2,   /* "Convolver: block size must be a multiple of 8!"
                                (apf/apf/convolver.h:163-166) */
  SSR_ERR_UNSUPPORTED = -3,  /* block size above 8192 (every multiple of 8 up to there is accepted) */
  SSR_ERR_CUDA = -4,         /* CUDA runtime failure / no device */
  SSR_ERR_NOMEM = -5,
  SSR_ERR_STATE = -6,        /* call order violated */
  SSR_ERR_FILE = -7          /* IR set shape mismatch (apf/apf/sndfiletools.h:48-87,
                                src/binauralrenderer.h:124-128, src/brsrenderer.h:101-105) */
} ssr_status;

typedef struct ssr_engine ssr_engine;
typedef struct ssr_filterset ssr_filterset;
typedef struct ssr_input ssr_input;
typedef struct ssr_output ssr_output;
typedef struct ssr_renderer ssr_renderer;
typedef struct ssr_gather ssr_gather;
typedef struct ssr_hostgather ssr_hostgather;

const char* ssr_last_error(void);
int ssr_abi_version(void);

/* ------------------------------------------------------------------ engine */

typedef struct ssr_engine_config
{
  int device;        /* CUDA device ordinal */
  int block_size;    /* B; reference parameter "block_size" (apf/apf/pointer_policy.h:75) */
  int sample_rate;   /* reference parameter "sample_rate" (pointer_policy.h:74) */
  void* stream;      /* cudaStream_t to run on, or NULL for an engine-owned stream */
  int mac_variant;   /* 0 = default (bulk-copy staged MAC for the four-row plans at
                        B >= 1024); other values select the other MAC kernel builds
                        (2 = register-staged; profiles/r1_mac_variants.md) */
  int reserved[7];
} ssr_engine_config;

/* replaces: apf::MimoProcessor construction for this path
 * (apf/apf/mimoprocessor.h:406-425) + per-object FFTW plan creation
 * (apf/apf/convolver.h:176-181, 421-423). */
int ssr_engine_create(const ssr_engine_config* cfg, ssr_engine** out);
void ssr_engine_destroy(ssr_engine* e);
int ssr_engine_block_size(const ssr_engine* e);
/* Blocks until all queued device work of this engine has finished. */
int ssr_engine_synchronize(ssr_engine* e);

typedef struct ssr_engine_stats
{
  /* accumulated since the last reset, device time from CUDA events (ms) */
  double ms_fft;          /* forward transforms of input blocks */
  double ms_mac;          /* spectral multiply-accumulate kernel */
  double ms_ifft;         /* inverse transform + crossfade epilogue */
  double ms_total;        /* first kernel start -> last kernel end */
  uint64_t blocks;        /* timed block steps */
  uint64_t mac_launches;  /* kernel launches, by kind */
  uint64_t fft_launches;
  uint64_t ifft_launches;
  uint64_t mac_bytes;     /* ALGORITHMIC bytes of the timed MAC launches:
                             8*B * (H partitions read + X partitions + accumulators written)
                             = SURVEY 8(d) formula */
  uint64_t mac_terms;     /* partition-MACs (8*B flops each at 8 flops / complex MAC) */
  uint64_t h2d_bytes, d2h_bytes;
  int timing_enabled;
  int timing_every;          /* every timing_every-th block step carries the stage events */
  uint64_t launches_total;   /* every kernel the engine launched (timed or not, graph replays included) */
  int last_mac_grid;         /* CTAs of the most recent MAC launch (ties ncu captures to the plan that ran) */
  int reserved[3];
} ssr_engine_stats;

/* Per-stage CUDA events.  enabled = 0: off; 1: every block step; k > 1: every k-th
 * block step (the four event records cost ~10 us of stream time per block, which
 * matters for the 50 us blocks of the small workloads; ms_*, blocks, mac_bytes and
 * the per-kind launch counts then describe the sampled steps only). */
int ssr_engine_set_timing(ssr_engine* e, int enabled);
int ssr_engine_get_stats(ssr_engine* e, ssr_engine_stats* out, int reset);
/* Development aid (engines that run the bulk-copy MAC, B >= 1024, else SSR_ERR_STATE): per
 * CTA of the most recent MAC launch {start, first data, end, busy time accumulated since the
 * last share calibration} in %globaltimer nanoseconds; out holds 4 * max_ctas values. */
int ssr_engine_debug_mac_trace(ssr_engine* e, uint64_t* out, int max_ctas);

/* ------------------------------------------------------------- filter sets */

/* A set of `channels` filters, each `partitions` partition spectra, resident in
 * HBM (north_star item 1).  Replaces apf::conv::Filter and containers of them
 * (apf/apf/convolver.h:100-117; hrtf_set_t src/binauralrenderer.h:72,
 * brtf_set_t src/brsrenderer.h:213).  A new set is all "zero"-flagged, like
 * Filter(block, partitions) (convolver.h:103-107). */
int ssr_filterset_create(ssr_engine* e, int channels, int partitions, ssr_filterset** out);
void ssr_filterset_destroy(ssr_filterset* fs);
int ssr_filterset_channels(const ssr_filterset* fs);
int ssr_filterset_partitions(const ssr_filterset* fs);

/* Transform `frames` taps of every channel: channel c, tap t is
 * data[t * frame_stride + c * channel_stride]  (libsndfile readf layout =
 * frame_stride = channels, channel_stride = 1, the apf::fixed_matrix "slices",
 * apf/apf/container.h:357-364).  Equals Transform::prepare_filter per channel
 * (apf/apf/convolver.h:190-231): chunks of <= B taps, zero-padded to 2B,
 * forward FFT; partitions beyond the taps stay zero-flagged.  `first_channel`
 * selects where in the set the `channels` channels of `data` are stored. */
int ssr_filterset_load_time(ssr_filterset* fs, int first_channel, int channels,
                            const float* data, long frames,
                            long frame_stride, long channel_stride);

/* = TransformBase::prepare_partition (convolver.h:209-231) for one partition. */
int ssr_filterset_set_partition_time(ssr_filterset* fs, int channel, int partition,
                                     const float* taps, int n);

/* Deterministic synthetic IRs generated ON THE DEVICE (no host copy), see
 * DESIGN.md "synthetic workload": tap t of channel c is
 * u(seed, c, t) * envelope[t] * scale, u uniform in [-1, 1).  The same generator
 * exists on the host as ssr_synth_ir() for parity tests. */
int ssr_filterset_generate(ssr_filterset* fs, int first_channel, int channels, long taps,
                           uint64_t seed, uint64_t channel_id_offset,
                           const float* envelope, float scale);
void ssr_synth_ir(uint64_t seed, uint64_t channel_id, long taps, const float* envelope,
                  float scale, float* out);
/* same generator for signal blocks: uniform [-1, 1) */
void ssr_synth_uniform(uint64_t seed, uint64_t stream_id, uint64_t first_index, long n, float* out);

/* dst[dch] = (1 - t) * a[ach] + t * b[bch], honouring zero flags and unequal
 * partition counts = apf::conv::transform_nested with the lerp of
 * src/binauralrenderer.h:372-377 (convolver.h:773-817). */
int ssr_filterset_lerp(ssr_filterset* dst, int dch, const ssr_filterset* a, int ach,
                       const ssr_filterset* b, int bch, float t);

/* Parity / debug: copies one partition spectrum to the host in the REFERENCE's
 * "sorted halfcomplex" layout (convolver.h:236-268), 2B floats; *zero receives
 * the zero flag. */
int ssr_filterset_download(ssr_filterset* fs, int channel, int partition,
                           float* sorted_out, int* zero);
/* The inverse: writes one partition spectrum given in the reference's layout, or
 * (zero != 0, sorted_in may be NULL) sets its zero flag.  Load-time tooling: the
 * C++ facade's transform_nested with an arbitrary host functor
 * (apf/apf/convolver.h:773-817) is download -> functor -> upload. */
int ssr_filterset_upload(ssr_filterset* fs, int channel, int partition,
                         const float* sorted_in, int zero);

/* ------------------------------------------- Level 1: apf::conv class API */

/* = apf::conv::Input(block, partitions) (convolver.h:296-318) */
int ssr_input_create(ssr_engine* e, int partitions, ssr_input** out);
void ssr_input_destroy(ssr_input* in);
int ssr_input_partitions(const ssr_input* in);
/* = Input::add_block (convolver.h:324-375): B host floats.  The block is copied
 * into the engine's pinned staging buffer (x may be reused on return); its forward
 * transform runs with the first ssr_output_convolve() on this input (one fused
 * launch: forward FFT + MAC + inverse), or at the latest with the next add_block /
 * download, so the delay line advances exactly once per add_block like the
 * reference's. */
int ssr_input_add_block(ssr_input* in, const float* x);
/* parity / debug: spectrum `p` blocks ago (0 = newest) in the reference layout */
int ssr_input_download(ssr_input* in, int p, float* sorted_out);

typedef enum ssr_output_kind { SSR_OUTPUT_DYNAMIC = 0, SSR_OUTPUT_STATIC = 1 } ssr_output_kind;

/* = apf::conv::Output(input) / StaticOutput(input, ...) (convolver.h:618-634,
 * 705-726).  Starts with every partition "empty" (convolver.h:415-417). */
int ssr_output_create(ssr_input* in, ssr_output_kind kind, ssr_output** out);
void ssr_output_destroy(ssr_output* o);
/* = StaticOutput::_set_filter (convolver.h:729-739) */
int ssr_output_bind_static(ssr_output* o, const ssr_filterset* fs, int channel);
/* = Output::set_filter / queues_empty / rotate_queues (convolver.h:642-699) */
int ssr_output_set_filter(ssr_output* o, const ssr_filterset* fs, int channel);
int ssr_output_queues_empty(const ssr_output* o);  /* 1 / 0 */
int ssr_output_rotate_queues(ssr_output* o);
/* = OutputBase::convolve(weight) (convolver.h:435-465).  *result points at B
 * floats (engine-owned pinned memory), valid until the next convolve() on `o`.
 * Term lists of up to 64 partitions run as ONE single-CTA launch that reads the
 * staged input block and writes the result block through mapped host memory
 * (the call returns when the block has landed); longer ones use the MAC and
 * inverse kernels of the batched path. */
int ssr_output_convolve(ssr_output* o, float weight, const float** result);

/* ------------------------- Level 2: renderer block step (the batched path) */

typedef enum ssr_renderer_kind
{
  SSR_RENDERER_GENERIC = 0,   /* src/genericrenderer.h */
  SSR_RENDERER_BINAURAL = 1,  /* src/binauralrenderer.h */
  SSR_RENDERER_BRS = 2,       /* src/brsrenderer.h */
  SSR_RENDERER_PREFILTER_BANK = 3  /* the per-input StaticConvolver stage of the WFS
                                      renderer (src/wfsrenderer.h:72-88, 104-123): N
                                      inputs -> N outputs, every input convolved with
                                      the shared pre-filter; weight 1, no crossfade */
} ssr_renderer_kind;

typedef struct ssr_renderer_config
{
  ssr_renderer_kind kind;
  int outputs;                       /* Generic: loudspeakers M (total); Binaural / BRS: must be 2 */
  int output_first, output_count;    /* Generic sharding by output channel:
                                        this engine renders outputs [first, first+count);
                                        count = 0 means all */
  float master_volume_correction_db; /* "master_volume_correction" (src/rendererbase.h:234-235) */
  int reserved[8];
} ssr_renderer_config;

int ssr_renderer_create(ssr_engine* e, const ssr_renderer_config* cfg, ssr_renderer** out);
void ssr_renderer_destroy(ssr_renderer* r);

/* Binaural: = BinauralRenderer::_load_hrtfs (src/binauralrenderer.h:117-169).
 * data is frames x channels interleaved (libsndfile readf); channels must be
 * even (else SSR_ERR_FILE "Number of channels in HRIR file must be a multiple
 * of 2!"); hrir_size = 0 means all frames. */
int ssr_renderer_load_hrirs(ssr_renderer* r, const float* data, long frames, int channels,
                            long hrir_size);
/* Prefilter bank: = WfsRenderer ctor (src/wfsrenderer.h:75-87): single-channel
 * pre-filter, `frames` taps; must be loaded before inputs are added.  Inputs are
 * added with ssr_renderer_add_source(r, NULL, 0, 0, &id); output i belongs to
 * input i (ssr_renderer_local_outputs() == number of inputs). */
int ssr_renderer_load_prefilter(ssr_renderer* r, const float* taps, long frames);
/* synthetic HRIR set generated on the device (bench workloads) */
int ssr_renderer_generate_hrirs(ssr_renderer* r, int channels, long taps, uint64_t seed,
                                const float* envelope, float scale);

/* = RendererBase::add_source (src/rendererbase.h:247-287) with the Source
 * constructors of the three renderers:
 *   Generic  (src/genericrenderer.h:88-120): ir = frames x M interleaved, channels
 *            must equal the renderer's TOTAL outputs (SSR_ERR_FILE otherwise);
 *   BRS      (src/brsrenderer.h:90-139): ir = frames x (2*angles);
 *   Binaural (src/binauralrenderer.h:235-259): ir must be NULL.
 * Returns the new source id in *id. */
int ssr_renderer_add_source(ssr_renderer* r, const float* ir, long frames, int channels, int* id);
/* synthetic IR set for a source generated on the device (Generic / BRS) */
int ssr_renderer_add_source_synthetic(ssr_renderer* r, long taps, int channels, uint64_t seed,
                                      uint64_t channel_id_offset, const float* envelope,
                                      float scale, int* id);
int ssr_renderer_rem_source(ssr_renderer* r, int id);
int ssr_renderer_sources(const ssr_renderer* r);

/* Control inputs.  THREAD-SAFE HAND-OFF: these setters may be called from any thread,
 * also while another thread is inside ssr_renderer_process / submit / wait (the GUI,
 * network and tracker threads of the reference, src/rendersubscriber.h:164-169).  Like
 * apf::SharedData::operator= (apf/apf/shareddata.h:57-60) they do not touch the
 * renderer: they queue a command (apf::CommandQueue::push, apf/apf/commandqueue.h:
 * 185-210; 4096 slots, several control threads are serialised internally), and the
 * audio thread executes everything queued, in order, at the start of its next block
 * (apf/apf/mimoprocessor.h:374).  A command for a source that has been removed
 * meanwhile is dropped.  SSR_ERR_STATE: the queue stayed full for 0.2 s (no thread
 * is rendering blocks).  Adding / removing sources, loading filters and the level
 * meters remain calls of the thread that owns the audio callback.
 *
 * per-source (src/rendererbase.h:419-423) */
int ssr_source_set_gain(ssr_renderer* r, int id, float gain);
int ssr_source_set_mute(ssr_renderer* r, int id, int mute);
int ssr_source_set_position(ssr_renderer* r, int id, float x, float y);
int ssr_source_set_model(ssr_renderer* r, int id, int plane);  /* 0 point, 1 plane */

/* global (RendererBase::State, src/rendererbase.h:84-103) */
int ssr_renderer_set_reference_position(ssr_renderer* r, float x, float y);
int ssr_renderer_set_reference_orientation(ssr_renderer* r, float azimuth_deg);
int ssr_renderer_set_reference_offset_position(ssr_renderer* r, float x, float y);
int ssr_renderer_set_reference_offset_orientation(ssr_renderer* r, float azimuth_deg);
int ssr_renderer_set_master_volume(ssr_renderer* r, float v);
int ssr_renderer_set_processing(ssr_renderer* r, int on);
int ssr_renderer_set_amplitude_reference_distance(ssr_renderer* r, float d);

/* The state the audio thread rendered its LAST block with (= SharedData::get() on the
 * audio side; queued commands are not visible before the block that applies them).
 * To be called by the thread that owns the audio callback. */
typedef struct ssr_renderer_state
{
  float reference_x, reference_y, reference_azimuth;
  float reference_offset_x, reference_offset_y, reference_offset_azimuth;
  float master_volume, amplitude_reference_distance;
  int processing;
  int pending_commands;      /* queued, not applied yet */
  uint64_t blocks;           /* blocks started so far */
  int reserved[4];
} ssr_renderer_state;
typedef struct ssr_source_state
{
  float gain; int mute; float x, y; int plane;
  float weighting_factor;    /* RendererBase::Source::weighting_factor of the last block */
  int reserved[4];
} ssr_source_state;
int ssr_renderer_get_state(ssr_renderer* r, ssr_renderer_state* out);
int ssr_source_get_state(ssr_renderer* r, int id, ssr_source_state* out);

/* Level meters = the reference's apf::enable_queries policy
 * (src/rendererbase.h:431-437, 468-474; shown by the GUI, src/controller.h:406-528):
 * per source the maximum |input sample| of the last block (pre-fader) and that
 * value times the source's weighting factor, per local output the maximum
 * |output sample|.  Computed inside the forward / inverse kernels when enabled;
 * off by default like the reference's disable_queries.  Arrays may be NULL. */
int ssr_renderer_set_level_metering(ssr_renderer* r, int on);
int ssr_renderer_get_levels(ssr_renderer* r, float* source_pre_fader, float* source_levels,
                            float* output_levels);

/* One audio block = apf::pointer_policy<float*>::audio_callback(n, in, out)
 * (apf/apf/pointer_policy.h:58,112-122): `in` has one pointer per source (in
 * add order), `out` one per LOCAL output; each points at B host floats.
 * H2D, forward FFT, MAC, inverse FFT + crossfade + source sum, D2H; returns
 * when `out` is filled. */
int ssr_renderer_process(ssr_renderer* r, int n, const float* const* in, float* const* out);

/* OFFLINE rendering of `nblocks` consecutive blocks in one call (the use case of
 * apf::mimoprocessor_file_io, apf/apf/mimoprocessor_file_io.h:40-117, where every
 * input block is known in advance): in[t * sources + i] / out[t * outputs + m] are
 * the channel pointers of block t.  Results are those of `nblocks` calls of
 * ssr_renderer_process.  Runs of 4 blocks during which no control changes (Generic
 * renderer, B >= 1024) are rendered by ONE pass over the filter spectra -- every H
 * tile is read once for 4 time steps -- which is not possible in real time, where
 * block t + 1 does not exist yet when block t is due.  Everything else
 * falls back to ssr_renderer_process block by block.
 * ssr_renderer_batched_blocks: how many blocks went through the batched pass. */
int ssr_renderer_process_batch(ssr_renderer* r, int n, int nblocks, const float* const* in, float* const* out);
int ssr_renderer_batched_blocks(const ssr_renderer* r, uint64_t* blocks);

/* Same step on DEVICE buffers, asynchronous on the engine's stream: d_in is
 * sources x B floats, d_out is local_outputs x B floats (both contiguous).
 * Used by the multi-GPU driver, which owns the buffers and the NCCL gather.
 * Calls queued back to back with nothing else on the stream in between overlap:
 * the forward transform of block t + 1 and the part of its multiply-accumulate
 * that does not need it run while block t's inverse transform still does
 * (results are bit-identical to one call at a time; every block must get its own
 * d_out if the caller reads them later, and d_in must stay valid until the block
 * has run).  Anything the caller queues on the stream between two calls -- a
 * copy, an event, a kernel -- simply keeps the blocks apart as before. */
int ssr_renderer_process_device(ssr_renderer* r, const float* d_in, float* d_out);

/* Pipelined host variant: queues the block and returns immediately; the result
 * of block t is delivered by ssr_renderer_wait().  Lets block t+1's H2D overlap
 * block t's compute (north_star item 5: CUDA-stream pipeline). */
int ssr_renderer_submit(ssr_renderer* r, int n, const float* const* in);
int ssr_renderer_wait(ssr_renderer* r, float* const* out);

int ssr_renderer_local_outputs(const ssr_renderer* r);

/* ------------------------------------------------------------------------
 * Multi-GPU output gather over NVLink peer memory.
 *
 * The Generic renderer shards by output channel: rank g renders outputs
 * [first_g, first_g + count_g) -- in the reference these are the Output items
 * apf::MimoProcessor spreads round-robin over its worker threads
 * (apf/apf/mimoprocessor.h:452-483, src/genericrenderer.h:122-180).  The one
 * exchange step is bringing the output blocks to the host-facing rank ("root").
 * With a gather bound, the inverse-FFT kernel of every rank stores its blocks
 * straight into the root's gather buffer (peer ranks: through NVLink peer
 * memory) and counts its CTAs in on a system-scope counter; the root runs a
 * one-warp wait kernel instead of a collective.  Two slots alternate by block
 * parity; a peer may run at most two blocks ahead of the root's release.
 *
 * One engine (= one GPU) per rank; ranks may be processes (CUDA IPC handle,
 * 64 bytes, carried by whatever the host uses: torch.distributed, MPI, a pipe)
 * or engines of one process (ssr_gather_connect_local).
 *   root:  ssr_gather_create -> ssr_gather_export -> (send handle)
 *   peer:  ssr_gather_create -> ssr_gather_import(handle)
 *   all:   ssr_renderer_bind_gather(renderer, gather)
 * Then ssr_renderer_process(r, n, in, out) is the same audio_callback on every
 * rank; on the root `out` has one pointer per output of the WHOLE renderer
 * (ssr_gather_total_outputs), on the other ranks `out` is ignored (may be NULL).
 * With ssr_renderer_process_device() (d_out ignored) the root calls
 * ssr_gather_wait() -> device pointer to total_outputs x B floats of this
 * block, valid until ssr_gather_release(), both stream-ordered; the root must
 * release every block (peers wait for it).
 *
 * Failure behaviour: every device-side wait gives up after 20 s
 * (SSR_B200_GATHER_TIMEOUT_MS overrides).  A peer that times out stores nothing;
 * the rank that timed out keeps a sticky flag, and from then on
 * ssr_renderer_process / ssr_renderer_wait on that rank return SSR_ERR_STATE
 * instead of delivering a stale block (ssr_gather_status() reports the same flag).
 *
 * One rank per device is the deployment shape.  Ranks that SHARE a device and
 * are driven by ONE host thread (single-GPU test rigs) must queue the root's
 * block last: the root's wait kernel spins on the device, and a peer's host
 * work that synchronises the device (a growing table) would otherwise wait for
 * it.  The host-direct path below (ssr_hostgather_*) never spins on a device
 * and has no such rule. */
#define SSR_GATHER_HANDLE_BYTES 64
int ssr_gather_create(ssr_engine* e, int world, int rank, int root, const int* outputs_per_rank,
                      ssr_gather** out);
void ssr_gather_destroy(ssr_gather* g);
int ssr_gather_export(ssr_gather* g, void* handle /* SSR_GATHER_HANDLE_BYTES */);
int ssr_gather_import(ssr_gather* g, const void* handle);
int ssr_gather_connect_local(ssr_gather* g, ssr_gather* root);
int ssr_renderer_bind_gather(ssr_renderer* r, ssr_gather* g /* NULL: unbind */);
int ssr_gather_wait(ssr_gather* g, const float** d_block);
int ssr_gather_release(ssr_gather* g);
int ssr_gather_status(ssr_gather* g);  /* SSR_OK, or SSR_ERR_STATE after a time-out (synchronises) */
int ssr_gather_total_outputs(const ssr_gather* g);

/* ------------------------------------------------------------------------
 * Host-direct output delivery for a sharded renderer (the HOST-buffer call).
 *
 * In the reference every Output item writes the caller's host channel itself
 * (apf/apf/pointer_policy.h:112-122, apf/apf/mimoprocessor.h:452-483).  The
 * same holds here across GPUs: all ranks map ONE page-locked POSIX shared-memory
 * segment ("/name", created by the root) and rank g's inverse kernel stores its
 * output blocks straight into it over that GPU's own PCIe link; the CTA that
 * finishes last publishes the rank's flag behind the data.  No GPU waits for
 * another GPU, nothing funnels through the root's device, and no D2H copy runs.
 *
 *   every rank: ssr_hostgather_create(e, world, rank, root, counts, "/name", &g)
 *               (peers wait up to the time-out for the root's segment to appear)
 *               ssr_renderer_bind_hostgather(renderer, g)
 *   per block:  ssr_renderer_process(r, n, in, out) on every rank.  On the root
 *               it returns once every rank's blocks of this block have landed;
 *               `out` has one pointer per output of the WHOLE renderer.  When
 *               out[m] == slot + m * B with slot = ssr_hostgather_slot(g,
 *               ssr_hostgather_next_slot(g)) the caller reads the blocks in place
 *               (no copy at all); any other pointers receive a host memcpy.  On
 *               the other ranks `out` is ignored (may be NULL).
 * Two slots alternate by block parity: a slot's blocks stay valid until the
 * root's second-next ssr_renderer_process.  A rank that waits longer than the
 * time-out (20 s, SSR_B200_GATHER_TIMEOUT_MS) fails its call with SSR_ERR_STATE.
 * ssr_renderer_process_device / submit / wait are not served by this path (they
 * use the peer-memory gather above). */
int ssr_hostgather_create(ssr_engine* e, int world, int rank, int root, const int* outputs_per_rank,
                          const char* shm_name, ssr_hostgather** out);
void ssr_hostgather_destroy(ssr_hostgather* g);
int ssr_renderer_bind_hostgather(ssr_renderer* r, ssr_hostgather* g /* NULL: unbind */);
int ssr_hostgather_total_outputs(const ssr_hostgather* g);
int ssr_hostgather_slot(ssr_hostgather* g, int slot /* 0 | 1 */, float** host_blocks);
int ssr_hostgather_next_slot(const ssr_hostgather* g);

/* ------------------------------------------------- non-uniform partitioning */
/* SURVEY 8(f) row 4 -- NOT in the reference (its manual only notes the trade-off of the
 * partition size, doc/manual/renderers.tex:199-204): a GenericRenderer whose impulse
 * responses are cut into LEVELS of growing partition size.  Level 0 has `level_partitions[0]`
 * partitions of the block size B and is the reference's algorithm unchanged
 * (apf/apf/convolver.h:324-375,435-465; src/genericrenderer.h:122-203); level l > 0 has
 * `level_partitions[l]` partitions of level_mult[l] * B samples, rendered by the same engine at
 * that block size every level_mult[l] blocks, its outputs spread over the blocks of a period
 * (csrc/nonuniform.cu).  Zero added latency, the same output within rounding for constant
 * source weights; a weight change reaches a tail level at that level's block rate.  The
 * filter bytes read per block fall from P to sum_l level_partitions[l] partitions' worth.
 * Rules (SSR_ERR_INVALID otherwise): level_mult[0] = 1; level_mult increasing powers of two
 * with level_mult[l] * B <= 8192; level l > 0 starts (sum of the lower levels' taps) at a
 * multiple of its partition size and at least two of its partitions into the response.
 * Sources are added before the first block. */
typedef struct ssr_nu ssr_nu;
typedef struct ssr_nu_config
{
  int device, block_size, sample_rate;
  int outputs;                       /* loudspeakers M (total) */
  int output_first, output_count;    /* output shard of this engine; count = 0: all */
  int levels;                        /* 1..4 */
  int level_mult[4];
  int level_partitions[4];
  float master_volume_correction_db;
  int reserved[8];
} ssr_nu_config;
int ssr_nu_create(const ssr_nu_config* cfg, ssr_nu** out);
void ssr_nu_destroy(ssr_nu* r);
/* as ssr_renderer_add_source / _synthetic for the Generic renderer; frames (taps) must not
 * exceed ssr_nu_taps_covered() */
int ssr_nu_add_source(ssr_nu* r, const float* ir, long frames, int channels, int* id);
int ssr_nu_add_source_synthetic(ssr_nu* r, long taps, int channels, uint64_t seed, uint64_t channel_id_offset,
                                const float* envelope, float scale, int* id);
int ssr_nu_source_set_gain(ssr_nu* r, int id, float gain);
int ssr_nu_source_set_mute(ssr_nu* r, int id, int mute);
int ssr_nu_set_master_volume(ssr_nu* r, float volume);
int ssr_nu_sources(const ssr_nu* r);
int ssr_nu_local_outputs(const ssr_nu* r);
long ssr_nu_taps_covered(const ssr_nu* r);
/* one block: host channel pointers, n = frames = the block size (= ssr_renderer_process,
 * apf/apf/pointer_policy.h:112-122) / device arrays [sources][B] -> [local outputs][B],
 * asynchronous on ssr_nu_stream() */
int ssr_nu_process(ssr_nu* r, int n, const float* const* in, float* const* out);
int ssr_nu_process_device(ssr_nu* r, const float* d_in, float* d_out);
int ssr_nu_synchronize(ssr_nu* r);
void* ssr_nu_stream(ssr_nu* r);
/* kernels launched so far (all levels' engines + the mixer) */
int ssr_nu_launches(ssr_nu* r, uint64_t* launches);
/* ALGORITHMIC bytes (SURVEY 8d formula per MAC launch, summed) of block `block_index` */
int ssr_nu_block_mac_bytes(ssr_nu* r, uint64_t block_index, uint64_t* bytes);

#ifdef __cplusplus
}
#endif

#endif /* SSR_B200_H */
