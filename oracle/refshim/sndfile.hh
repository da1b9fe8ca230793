/* sndfile.hh SHIM -- test infrastructure only (oracle/).
 * libsndfile is absent from this image.  The reference only uses
 * SndfileHandle{ctor(name, SFM_READ), channels(), samplerate(), frames(),
 * readf(float*, n)} on the convolution path (apf/apf/sndfiletools.h:43-90,
 * src/binauralrenderer.h:119-140, src/brsrenderer.h:95-113,
 * src/genericrenderer.h:94-104).  This shim serves "files" from an in-memory
 * registry filled by the harness (name -> interleaved frames x channels).
 */
#ifndef SSR_B200_ORACLE_SNDFILE_SHIM_HH
#define SSR_B200_ORACLE_SNDFILE_SHIM_HH

#include <algorithm>
#include <cstdint>
#include <map>
#include <memory>
#include <string>
#include <vector>

typedef int64_t sf_count_t;
enum { SFM_READ = 0x10, SFM_WRITE = 0x20, SFM_RDWR = 0x30 };

namespace ssr_b200_sndshim
{
struct MemFile
{
  int channels = 0;
  int samplerate = 0;
  sf_count_t frames = 0;
  // either owned or borrowed interleaved data (frames x channels)
  std::vector<float> owned;
  const float* data = nullptr;
};

inline std::map<std::string, std::shared_ptr<MemFile>>& registry()
{
  static std::map<std::string, std::shared_ptr<MemFile>> r;
  return r;
}
}  // namespace ssr_b200_sndshim

class SndfileHandle
{
  public:
    SndfileHandle() {}
    SndfileHandle(const std::string& path, int mode = SFM_READ, int = 0
        , int = 0, int = 0)
    {
      if (mode != SFM_READ) return;
      auto& r = ssr_b200_sndshim::registry();
      auto it = r.find(path);
      if (it != r.end()) _f = it->second;
    }
    int channels() const { return _f ? _f->channels : 0; }
    int samplerate() const { return _f ? _f->samplerate : 0; }
    sf_count_t frames() const { return _f ? _f->frames : 0; }
    sf_count_t readf(float* ptr, sf_count_t n)
    {
      if (!_f) return 0;
      sf_count_t avail = std::max<sf_count_t>(0, _f->frames - _pos);
      n = std::min(n, avail);
      const float* src = _f->data + _pos * _f->channels;
      std::copy(src, src + n * _f->channels, ptr);
      _pos += n;
      return n;
    }
  private:
    std::shared_ptr<ssr_b200_sndshim::MemFile> _f;
    sf_count_t _pos = 0;
};

#endif
