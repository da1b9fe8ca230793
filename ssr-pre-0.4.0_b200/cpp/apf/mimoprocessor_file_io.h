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
// The file is streamed block by block (WavReader / WavWriter), as the reference streams
// it through libsndfile.
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

  // the file is STREAMED block by block like the reference does with
  // SndfileHandle::readf / writef (apf/apf/mimoprocessor_file_io.h:80-110): memory is
  // a few blocks, whatever the length of the recording
  ssr::b200::WavReader in;
  if (infilename == "" || !in.open(infilename) || !in.channels)
  {
    std::cout << "apf::load_sndfile(): \"" << infilename << "\" couldn't be loaded!" << std::endl;
    return 1;
  }

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
  const long frames = in.frames;
  const long blocks = (frames + blocksize - 1) / blocksize;

  ssr::b200::WavWriter out;
  out.open(outfilename, nout, in.samplerate, in.format, in.bits);

  // interleaved file blocks <-> de-interleaved channel blocks (the reference uses
  // fixed_matrix transposes)
  std::vector<float> ileave_in(size_t(blocksize) * nin), ileave_out(size_t(blocksize) * nout);
  auto read_block = [&](float* const* channel_ptrs)
  {
    const long n = in.read(ileave_in.data(), blocksize);
    for (int c = 0; c < nin; ++c)
    {
      float* dst = channel_ptrs[c];
      for (long i = 0; i < n; ++i) dst[i] = ileave_in[size_t(i) * nin + c];
      for (long i = n; i < blocksize; ++i) dst[i] = 0.0f;   // last, partial block: zero-padded
    }
  };
  auto write_block = [&](long b, float* const* channel_ptrs)
  {
    const long n = std::min<long>(blocksize, frames - b * blocksize);
    for (int c = 0; c < nout; ++c)
      for (long i = 0; i < n; ++i) ileave_out[size_t(i) * nout + c] = channel_ptrs[c][i];
    out.write(ileave_out.data(), n);
  };

  processor.activate();
  const auto t0 = std::chrono::steady_clock::now();

  if (processor.batchable())
  {
    // Every block is known in advance: hand the engine chunks of blocks, so that runs of
    // constant-control blocks share one pass over the filter spectra
    // (ssr_renderer_process_batch; same results as block by block).
    const long chunk = 16;
    std::vector<float> ibuf(size_t(chunk) * nin * blocksize), obuf(size_t(chunk) * nout * blocksize);
    std::vector<float*> ip(size_t(chunk) * nin), op(size_t(chunk) * nout);
    std::vector<const float*> ipc(ip.size());
    for (size_t i = 0; i < ip.size(); ++i) { ip[i] = ibuf.data() + i * blocksize; ipc[i] = ip[i]; }
    for (size_t i = 0; i < op.size(); ++i) op[i] = obuf.data() + i * blocksize;
    for (long b0 = 0; b0 < blocks; b0 += chunk)
    {
      const long nb = std::min<long>(chunk, blocks - b0);
      for (long k = 0; k < nb; ++k) read_block(ip.data() + size_t(k) * nin);
      processor.audio_callback_batch(int(blocksize), int(nb), ipc.data(), op.data());
      for (long k = 0; k < nb; ++k) write_block(b0 + k, op.data() + size_t(k) * nout);
    }
  }
  else
  {
    // two blocks in flight: block b + 1 is read, copied and queued while block b computes
    // (through the processor, so that a renderer sharded over several devices works too)
    std::vector<float> bin(size_t(nin) * blocksize), bout(size_t(nout) * blocksize);
    std::vector<float*> in_ptrs(nin), out_ptrs(nout);
    for (int c = 0; c < nin; ++c) in_ptrs[c] = bin.data() + size_t(c) * blocksize;
    for (int c = 0; c < nout; ++c) out_ptrs[c] = bout.data() + size_t(c) * blocksize;
    for (long b = 0; b < blocks; ++b)
    {
      read_block(in_ptrs.data());
      processor.submit(int(blocksize), in_ptrs.data());
      if (b > 0)
      {
        processor.wait(out_ptrs.data());
        write_block(b - 1, out_ptrs.data());
      }
    }
    if (blocks > 0)
    {
      processor.wait(out_ptrs.data());
      write_block(blocks - 1, out_ptrs.data());
    }
  }

  const double secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  std::cout << "processing: " << secs << " s ("
    << (secs > 0 ? double(frames) / in.samplerate / secs : 0.0) << " x realtime)" << std::endl;

  processor.deactivate();
  out.close();
  return 0;
}

}  // namespace apf

#endif
