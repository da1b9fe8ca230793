// apf/mimoprocessor_file_io.h -- offline driver: run a renderer over a
// multichannel sound file.  Same entry point as the reference
// (apf/apf/mimoprocessor_file_io.h:40-117):
//
//   int apf::mimoprocessor_file_io(processor, infilename, outfilename);
//
// `processor` is any object with sample_rate(), in_channels(), out_channels(),
// block_size(), activate(), audio_callback(n, in, out), deactivate() -- in
// particular the GPU renderer adaptors of ssr/renderers.h.  WAV files are read
// and written by the small reader/writer of ssr/renderers.h instead of
// libsndfile; the output uses the input's sample format.
//
// Differences from the reference: the same return codes for sample-rate (42) and
// channel (666) mismatches; the last, partial block is zero-padded (the reference
// leaves the tail of its de-interleaving matrix from the previous block in place);
// blocks are pipelined two deep through the processor's submit() / wait()
// (ssr_renderer_submit / ssr_renderer_wait on every output shard; north_star
// item 5), which does not change the result.
#ifndef SSR_B200_APF_MIMOPROCESSOR_FILE_IO_H
#define SSR_B200_APF_MIMOPROCESSOR_FILE_IO_H

#include <chrono>
#include <iostream>
#include <string>
#include <vector>

#include "ssr/renderers.h"

namespace apf
{

template<typename Processor>
int mimoprocessor_file_io(Processor& processor
    , const std::string& infilename
    , const std::string& outfilename)
{
  std::cout << "Opening file \"" << infilename << "\" ..." << std::endl;

  ssr::b200::SoundFile in;
  try { in = ssr::b200::load_sndfile(infilename, 0, 0); }
  catch (const std::exception& e) { std::cout << e.what() << std::endl; return 1; }

  if (in.samplerate != static_cast<int>(processor.sample_rate()))
  {
    std::cout << "Samplerate mismatch!" << std::endl;
    return 42;
  }
  if (in.channels != processor.in_channels())
  {
    std::cout << "Input channel mismatch!" << std::endl;
    return 666;
  }

  std::cout << "frames: " << in.frames
    << " (" << in.frames / in.samplerate << " seconds)" << std::endl;
  std::cout << "channels: " << in.channels << std::endl;
  std::cout << "samplerate: " << in.samplerate << std::endl;

  const int blocksize = processor.block_size();
  const int nin = in.channels, nout = processor.out_channels();
  const long blocks = (in.frames + blocksize - 1) / blocksize;

  // de-interleaved block buffers (the reference uses fixed_matrix transposes)
  std::vector<float> bin(size_t(nin) * blocksize), bout(size_t(nout) * blocksize);
  std::vector<float*> in_ptrs(nin), out_ptrs(nout);
  for (int c = 0; c < nin; ++c) in_ptrs[c] = bin.data() + size_t(c) * blocksize;
  for (int c = 0; c < nout; ++c) out_ptrs[c] = bout.data() + size_t(c) * blocksize;
  std::vector<float> result(size_t(in.frames) * nout);

  processor.activate();
  const auto t0 = std::chrono::steady_clock::now();

  auto load_block = [&](long b)
  {
    const long first = b * blocksize;
    const long n = std::min<long>(blocksize, in.frames - first);
    for (int c = 0; c < nin; ++c)
    {
      float* dst = in_ptrs[c];
      for (long i = 0; i < n; ++i) dst[i] = in.data[size_t(first + i) * nin + c];
      for (long i = n; i < blocksize; ++i) dst[i] = 0.0f;
    }
  };
  auto store_block = [&](long b)
  {
    const long first = b * blocksize;
    const long n = std::min<long>(blocksize, in.frames - first);
    for (int c = 0; c < nout; ++c)
      for (long i = 0; i < n; ++i) result[size_t(first + i) * nout + c] = out_ptrs[c][i];
  };

  if (processor.batchable())
  {
    // Every block is known in advance: hand the engine chunks of blocks, so that runs of
    // constant-control blocks share one pass over the filter spectra
    // (ssr_renderer_process_batch; same results as block by block).
    const long chunk = 16;
    std::vector<std::vector<float>> ibuf(size_t(chunk) * nin, std::vector<float>(size_t(blocksize)));
    std::vector<std::vector<float>> obuf(size_t(chunk) * nout, std::vector<float>(size_t(blocksize)));
    std::vector<const float*> ip(ibuf.size());
    std::vector<float*> op(obuf.size());
    for (size_t i = 0; i < ibuf.size(); ++i) ip[i] = ibuf[i].data();
    for (size_t i = 0; i < obuf.size(); ++i) op[i] = obuf[i].data();
    for (long b0 = 0; b0 < blocks; b0 += chunk)
    {
      const long nb = std::min<long>(chunk, blocks - b0);
      for (long k = 0; k < nb; ++k)
      {
        const long first = (b0 + k) * blocksize;
        const long n = std::min<long>(blocksize, in.frames - first);
        for (int c = 0; c < nin; ++c)
        {
          float* dst = ibuf[size_t(k) * nin + c].data();
          for (long i = 0; i < n; ++i) dst[i] = in.data[size_t(first + i) * nin + c];
          for (long i = n; i < blocksize; ++i) dst[i] = 0.0f;
        }
      }
      processor.audio_callback_batch(int(blocksize), int(nb), ip.data(), op.data());
      for (long k = 0; k < nb; ++k)
      {
        const long first = (b0 + k) * blocksize;
        const long n = std::min<long>(blocksize, in.frames - first);
        for (int c = 0; c < nout; ++c)
          for (long i = 0; i < n; ++i) result[size_t(first + i) * nout + c] = obuf[size_t(k) * nout + c][i];
      }
    }
  }
  else
  {
  // two blocks in flight: block b + 1 is copied and queued while block b computes
  // (through the processor, so that a renderer sharded over several devices works too)
  for (long b = 0; b < blocks; ++b)
  {
    load_block(b);
    processor.submit(int(blocksize), in_ptrs.data());
    if (b > 0)
    {
      processor.wait(out_ptrs.data());
      store_block(b - 1);
    }
  }
  if (blocks > 0)
  {
    processor.wait(out_ptrs.data());
    store_block(blocks - 1);
  }

  }

  const double secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  std::cout << "processing: " << secs << " s ("
    << (secs > 0 ? double(in.frames) / in.samplerate / secs : 0.0) << " x realtime)" << std::endl;

  processor.deactivate();

  ssr::b200::write_wav(outfilename, result.data(), in.frames, nout, in.samplerate, in.format, in.bits);
  return 0;
}

}  // namespace apf

#endif
