// apf/convolver.h -- drop-in C++ facade of the reference's apf::conv namespace
// (apf/apf/convolver.h:55-819) on top of the B200 engine's C ABI
// (include/ssr_b200.h).  Same class names, constructor signatures, member
// functions and exception behaviour; the arithmetic runs in CUDA kernels and the
// spectra live in HBM.  Header-only; link with -lssr_b200.
//
//   namespace apf::conv { min_partitions, Filter, Transform, Input, Output,
//                         StaticOutput, Convolver, StaticConvolver,
//                         transform_nested }
//
// Deviations from the reference (all documented in INTEGRATION.md):
//  * Filter does not expose host fft_node storage (spectra are device resident);
//    Filter::download() returns a partition in the reference's sorted layout.
//  * Transform::prepare_partition takes (first, last, filter, index) instead of
//    a fft_node&.
//  * transform_nested with apf::conv::lerp(t) -- the one functor the renderers
//    use (src/binauralrenderer.h:372-377) -- runs as a kernel; any other binary
//    host callable works too, through a host round trip of the spectra (it is
//    host code; a load-time operation in the reference as well).
//  * block sizes: every multiple of 8 up to 8192 (the reference has no upper
//    bound); larger ones throw std::logic_error with the engine's message.
//  * One engine per (device, block size) is created lazily; the device ordinal
//    comes from the environment variable SSR_B200_DEVICE (default 0).
#ifndef SSR_B200_APF_CONVOLVER_H
#define SSR_B200_APF_CONVOLVER_H

#include <cstdlib>
#include <iterator>
#include <map>
#include <memory>
#include <mutex>
#include <stdexcept>
#include <string>
#include <vector>

#include "ssr_b200.h"

namespace apf
{
namespace conv
{

namespace detail
{
inline void check(int rc)
{
  if (rc == SSR_OK) return;
  const std::string msg = ssr_last_error();
  // the reference throws std::logic_error for a bad block size
  // (convolver.h:163-166); argument errors are asserts there
  if (rc == SSR_ERR_BLOCK_SIZE || rc == SSR_ERR_UNSUPPORTED || rc == SSR_ERR_INVALID
      || rc == SSR_ERR_FILE || rc == SSR_ERR_STATE)
    throw std::logic_error(msg);
  throw std::runtime_error(msg);
}

struct EngineDeleter { void operator()(ssr_engine* e) const { ssr_engine_destroy(e); } };

/// Process-wide engine for a block size (FFT plans in the reference are also
/// shared per size through FFTW wisdom, convolver.h:169-175).
inline ssr_engine* engine_for(size_t block_size)
{
  static std::mutex m;
  static std::map<size_t, std::unique_ptr<ssr_engine, EngineDeleter>> engines;
  std::lock_guard<std::mutex> guard(m);
  auto it = engines.find(block_size);
  if (it != engines.end()) return it->second.get();
  ssr_engine_config cfg = ssr_engine_config();
  const char* dev = std::getenv("SSR_B200_DEVICE");
  cfg.device = dev ? std::atoi(dev) : 0;
  cfg.block_size = int(block_size);
  cfg.sample_rate = 0;
  ssr_engine* e = nullptr;
  check(ssr_engine_create(&cfg, &e));
  engines[block_size].reset(e);
  return e;
}

template<typename In>
std::vector<float> to_vector(In first, In last)
{
  std::vector<float> v;
  for (; first != last; ++first) v.push_back(static_cast<float>(*first));
  return v;
}
}  // namespace detail

/// Calculate necessary number of partitions for a given filter length
/// (convolver.h:59-62)
static inline size_t min_partitions(size_t block_size, size_t filter_size)
{
  return (filter_size + block_size - 1) / block_size;
}

/// Container holding a number of FFT blocks (convolver.h:100-117), resident in HBM.
struct Filter
{
  /// Constructor; create empty filter.
  Filter(size_t block_size_, size_t partitions_)
    : _block_size(block_size_), _partitions(partitions_), _set(nullptr)
  {
    if (partitions_ == 0) throw std::logic_error("Filter: partitions must be > 0");
    detail::check(ssr_filterset_create(detail::engine_for(block_size_), 1
          , int(partitions_), &_set));
  }

  /// Constructor from time domain coefficients (convolver.h:282-291).
  template<typename In>
  Filter(size_t block_size_, In first, In last, size_t partitions_ = 0)
    : _block_size(block_size_)
    , _partitions(partitions_ ? partitions_
        : min_partitions(block_size_, size_t(std::distance(first, last))))
    , _set(nullptr)
  {
    if (_partitions == 0) throw std::logic_error("Filter: partitions must be > 0");
    detail::check(ssr_filterset_create(detail::engine_for(block_size_), 1
          , int(_partitions), &_set));
    load(first, last);
  }

  Filter(const Filter&) = delete;
  Filter& operator=(const Filter&) = delete;
  Filter(Filter&& other) noexcept
    : _block_size(other._block_size), _partitions(other._partitions), _set(other._set)
  { other._set = nullptr; }
  ~Filter() { if (_set) ssr_filterset_destroy(_set); }

  size_t block_size() const { return _block_size; }
  size_t partition_size() const { return 2 * _block_size; }
  size_t partitions() const { return _partitions; }

  /// (= Transform::prepare_filter) transform time-domain samples into this filter
  template<typename In>
  void load(In first, In last)
  {
    auto taps = detail::to_vector(first, last);
    float dummy = 0.0f;
    detail::check(ssr_filterset_load_time(_set, 0, 1, taps.empty() ? &dummy : taps.data()
          , long(taps.size()), 1, 1));
  }

  /// Partition spectrum in the reference's sorted-halfcomplex layout
  /// (convolver.h:236-268); returns the zero flag.
  bool download(size_t partition, float* sorted_out) const
  {
    int zero = 0;
    detail::check(ssr_filterset_download(_set, 0, int(partition), sorted_out, &zero));
    return zero != 0;
  }

  ssr_filterset* handle() const { return _set; }

  private:
    size_t _block_size, _partitions;
    ssr_filterset* _set;
};

/// Forward-FFT-related functions (convolver.h:120-157, 271-280)
class TransformBase
{
  public:
    template<typename In>
    void prepare_filter(In first, In last, Filter& filter) const
    {
      filter.load(first, last);
    }

    /// FFT of one block into partition @p index of @p filter; returns the
    /// iterator to the first coefficient of the next block (convolver.h:209-231).
    template<typename In>
    In prepare_partition(In first, In last, Filter& filter, size_t index) const
    {
      auto chunk = std::min(_block_size, size_t(std::distance(first, last)));
      In stop = first;
      std::advance(stop, chunk);
      auto taps = detail::to_vector(first, stop);
      float dummy = 0.0f;
      detail::check(ssr_filterset_set_partition_time(filter.handle(), 0, int(index)
            , taps.empty() ? &dummy : taps.data(), int(taps.size())));
      return stop;
    }

    size_t block_size() const { return _block_size; }
    size_t partition_size() const { return _partition_size; }

  protected:
    explicit TransformBase(size_t block_size_)
      : _block_size(block_size_), _partition_size(2 * block_size_)
    {
      if (_block_size % 8 != 0)
      {
        throw std::logic_error("Convolver: block size must be a multiple of 8!");
      }
      (void)detail::engine_for(block_size_);  // throws for unsupported sizes / no device
    }

  private:
    const size_t _block_size;
    const size_t _partition_size;
};

/// Helper class to prepare filters
struct Transform : TransformBase
{
  Transform(size_t block_size_) : TransformBase(block_size_) {}
};

/// Input stage of convolution (convolver.h:296-375)
struct Input : TransformBase
{
  Input(size_t block_size_, size_t partitions_)
    : TransformBase(block_size_), _partitions(partitions_), _in(nullptr)
  {
    if (partitions_ == 0) throw std::logic_error("Input: partitions must be > 0");
    detail::check(ssr_input_create(detail::engine_for(block_size_), int(partitions_), &_in));
  }
  Input(const Input&) = delete;
  Input& operator=(const Input&) = delete;
  /// Movable like the reference's Input (convolver.h:135); Outputs created from
  /// an Input keep a reference to it, so do not move an Input that has Outputs.
  Input(Input&& other) noexcept
    : TransformBase(other), _partitions(other._partitions), _in(other._in)
  { other._in = nullptr; }
  ~Input() { if (_in) ssr_input_destroy(_in); }

  /// Add a block of time-domain input samples (any forward iterator).
  template<typename In>
  void add_block(In first)
  {
    _tmp.resize(this->block_size());
    for (auto& x : _tmp) { x = static_cast<float>(*first); ++first; }
    detail::check(ssr_input_add_block(_in, _tmp.data()));
  }

  size_t partitions() const { return _partitions; }
  ssr_input* handle() const { return _in; }

  private:
    size_t _partitions;
    ssr_input* _in;
    std::vector<float> _tmp;
};

/// Base class for Output and StaticOutput (convolver.h:378-465)
class OutputBase
{
  public:
    /// Fast convolution of one audio block; returns a pointer to block_size()
    /// samples (engine-owned), valid until the next convolve() on this object.
    float* convolve(float weight = 1.0f)
    {
      const float* res = nullptr;
      detail::check(ssr_output_convolve(_out, weight, &res));
      return const_cast<float*>(res);
    }

    size_t block_size() const { return _input.block_size(); }
    size_t partitions() const { return _input.partitions(); }

    OutputBase(const OutputBase&) = delete;
    OutputBase& operator=(const OutputBase&) = delete;
    OutputBase(OutputBase&& other) noexcept : _input(other._input), _out(other._out)
    { other._out = nullptr; }

  protected:
    OutputBase(const Input& input, ssr_output_kind kind) : _input(input), _out(nullptr)
    {
      detail::check(ssr_output_create(input.handle(), kind, &_out));
    }
    ~OutputBase() { if (_out) ssr_output_destroy(_out); }

    const Input& _input;
    ssr_output* _out;
};

/// Convolution engine (output part) with switchable filter (convolver.h:618-699)
class Output : public OutputBase
{
  public:
    Output(const Input& input) : OutputBase(input, SSR_OUTPUT_DYNAMIC) {}
    Output(Output&&) = default;

    /// The first partition is updated immediately, later ones with rotate_queues().
    /// @attention as in the reference, @p filter must outlive its use here.
    void set_filter(const Filter& filter)
    {
      detail::check(ssr_output_set_filter(_out, filter.handle(), 0));
    }
    bool queues_empty() const { return ssr_output_queues_empty(_out) != 0; }
    void rotate_queues() { detail::check(ssr_output_rotate_queues(_out)); }
};

/// Convolver output stage with static filter (convolver.h:705-743)
class StaticOutput : public OutputBase
{
  public:
    /// Constructor from time domain samples
    template<typename In>
    StaticOutput(const Input& input, In first, In last)
      : OutputBase(input, SSR_OUTPUT_STATIC)
    {
      _filter.reset(new Filter(input.block_size(), first, last, input.partitions()));
      detail::check(ssr_output_bind_static(_out, _filter->handle(), 0));
    }

    /// Constructor from existing frequency domain filter coefficients.
    /// @attention The filter coefficients are not copied, their lifetime must
    ///   exceed that of the StaticOutput!
    StaticOutput(const Input& input, const Filter& filter)
      : OutputBase(input, SSR_OUTPUT_STATIC)
    {
      detail::check(ssr_output_bind_static(_out, filter.handle(), 0));
    }

    StaticOutput(StaticOutput&&) = default;

  private:
    std::unique_ptr<Filter> _filter;
};

/// Combination of Input and Output
struct Convolver : Input, Output
{
  Convolver(size_t block_size_, size_t partitions_)
    : Input(block_size_, partitions_)
    , Output(*static_cast<Input*>(this))
  {}
  Convolver(Convolver&&) = default;
  using Input::block_size;
  using Input::partitions;
};

/// Combination of Input and StaticOutput
struct StaticConvolver : Input, StaticOutput
{
  template<typename In>
  StaticConvolver(size_t block_size_, In first, In last, size_t partitions_ = 0)
    : Input(block_size_, partitions_ ? partitions_
        : min_partitions(block_size_, size_t(std::distance(first, last))))
    , StaticOutput(*this, first, last)
  {}

  StaticConvolver(const Filter& filter, size_t partitions_ = 0)
    : Input(filter.block_size(), partitions_ ? partitions_ : filter.partitions())
    , StaticOutput(*this, filter)
  {}
  StaticConvolver(StaticConvolver&&) = default;
  using Input::block_size;
  using Input::partitions;
};

/// The binary function the renderers pass to transform_nested
/// (src/binauralrenderer.h:372-377): (1 - t) * one + t * two
struct lerp
{
  explicit lerp(float t_) : t(t_) {}
  float operator()(float one, float two) const { return (1.0f - t) * one + t * two; }
  float t;
};

/// transform_nested (convolver.h:773-817) for the lerp functor: one kernel
inline void transform_nested(const Filter& in1, const Filter& in2, Filter& out, lerp f)
{
  detail::check(ssr_filterset_lerp(out.handle(), 0, in1.handle(), 0, in2.handle(), 0, f.t));
}

/// transform_nested (convolver.h:773-817) for any binary host callable
/// float(float, float): out[p] = f(in1[p], in2[p]) element by element, with the
/// reference's zero-flag rules -- both inputs zero-flagged (or past their end) ->
/// result zero-flagged; one of them -> f(x, 0) resp. f(0, x).  The spectra make a
/// host round trip in the reference's layout (f is host code).
template<typename BinaryFunction>
void transform_nested(const Filter& in1, const Filter& in2, Filter& out, BinaryFunction f)
{
  if (in1.block_size() != out.block_size() || in2.block_size() != out.block_size())
    throw std::logic_error("transform_nested: block sizes differ");
  const size_t n = out.partition_size();
  std::vector<float> a(n), b(n), r(n);
  for (size_t p = 0; p < out.partitions(); ++p)
  {
    const bool zero1 = p >= in1.partitions() || in1.download(p, a.data());
    const bool zero2 = p >= in2.partitions() || in2.download(p, b.data());
    if (zero1 && zero2)
    {
      detail::check(ssr_filterset_upload(out.handle(), 0, int(p), nullptr, 1));
      continue;
    }
    for (size_t i = 0; i < n; ++i) r[i] = f(zero1 ? 0.0f : a[i], zero2 ? 0.0f : b[i]);
    detail::check(ssr_filterset_upload(out.handle(), 0, int(p), r.data(), 0));
  }
}

}  // namespace conv
}  // namespace apf

#endif
