// Host-memory entry points: H2D copy, device step, D2H copy, pipelined over chunks with three streams
// so that the PCIe transfers of neighbouring chunks overlap the kernels.  Chunks are consecutive
// pieces of the call; the per-object state is carried from chunk to chunk exactly as it is from call
// to call (every path is block-partition independent), so the result is the one-shot one.
#pragma once
#include <cstdlib>
#include "common.cuh"

#include <algorithm>

namespace tsdgpu {

HostStage &host_stage();
int host_stage_reserve(size_t in_bytes, size_t out_bytes);
// 2-D copies between caller memory and a device slot (runtime.cu).  Pinned / registered caller memory: one
// cudaMemcpy2DAsync on the copy stream.  Pageable memory: through the slot's pinned bounce buffer, filled (stage_in) or
// drained (stage_out: deferred until the slot is used again, or stage_flush) by the copy-thread pool.
int stage_in(int slot, void *dst_dev, size_t dpitch, const void *src_host, size_t spitch, size_t width, size_t height);
int stage_out(int slot, void *dst_host, size_t dpitch, const void *src_dev, size_t spitch, size_t width, size_t height);
int stage_flush();

// unit = one chunkable item (a sample position for the streaming filters, a transform for the FFT).
// copy_in(slot, first, count) / copy_out(slot, out_first, out_count) issue the 2-D copies,
// out_count_of(count) tells how many output units the NEXT step of `count` input units will produce,
// step(slot, count, &out_count) enqueues the device work on rt().stream.
template<class CopyIn, class OutCount, class Step, class CopyOut>
int host_pipeline(long long total, long long chunk, CopyIn copy_in, OutCount out_count_of, Step step, CopyOut copy_out,
                  long long *out_total)
{
  Runtime &r = rt();
  HostStage &hs = host_stage();
  hs.pend[0].active = hs.pend[1].active = false;   // nothing survives a call (an earlier call may have stopped on an error)
  long long produced = 0;
  int k = 0;
  for(long long first = 0; first < total; first += chunk, k++)
  {
    const int slot = k & 1;
    const long long count = std::min(chunk, total - first);
    // the input slot is free once the kernel of chunk k-2 has run; the output slot once its D2H is done
    TSD_CUDA(cudaStreamWaitEvent(r.copy_in, hs.ev_done[slot], 0));
    if(copy_in(slot, first, count)) return 1;
    TSD_CUDA(cudaEventRecord(hs.ev_in[slot], r.copy_in));
    TSD_CUDA(cudaStreamWaitEvent(r.stream, hs.ev_in[slot], 0));
    TSD_CUDA(cudaStreamWaitEvent(r.stream, hs.ev_out[slot], 0));
    const long long expect = out_count_of(count);
    long long got = 0;
    if(step(slot, count, &got)) return 1;
    if(got != expect) return fail("host pipeline: output count mismatch");
    TSD_CUDA(cudaEventRecord(hs.ev_done[slot], r.stream));
    if(got > 0)
    {
      TSD_CUDA(cudaStreamWaitEvent(r.copy_out, hs.ev_done[slot], 0));
      if(copy_out(slot, produced, got)) return 1;
    }
    TSD_CUDA(cudaEventRecord(hs.ev_out[slot], r.copy_out));
    produced += got;
  }
  TSD_CUDA(cudaStreamSynchronize(r.copy_out));
  TSD_CUDA(cudaStreamSynchronize(r.stream));
  if(stage_flush()) return 1;
  if(out_total) *out_total = produced;
  return 0;
}

// chunk length (in samples) for a batch of nchan channels: ~48 MiB of input per chunk
inline long long host_chunk_len(int nchan, size_t elem_bytes, long long n, long long align = 1)
{
  static const unsigned long long mb = [] {
    const char *v = getenv("TSDGPU_HOST_CHUNK_MB");   // tuning knob: input bytes per pipeline chunk
    return (unsigned long long) std::max(1, v ? atoi(v) : 48);
  }();
  long long c = (long long) ((mb << 20) / ((size_t) nchan * elem_bytes));
  c = std::max<long long>(c, 4096);
  c = std::max<long long>(align, (c / align) * align);
  return std::min(c, std::max<long long>(n, 1));
}

} // namespace tsdgpu
