// Host side of the carried collectives: the job queue, the per-GEMM byte budget, the stand-alone flush kernel and the
// "done" wait kernel.  Device role: carry.cuh.  Reference parity: torchacc/dist/fsdp.py:196-230.
#include "carry.h"

#include <stdlib.h>
#include <string.h>

#include <deque>
#include <mutex>

#include "carry.cuh"

namespace tb {

namespace {

struct Job {
  long long id;
  CarrySlice proto;
  uint32_t total, next;      // chunk space / first chunk not yet handed out
  double bytes_per_chunk;    // link bytes one chunk moves
};

std::mutex g_mu;             // forward runs on the main thread, backward on autograd's device thread
// Two FIFO queues.  [0] foreground: jobs somebody will wait for soon (the next unit's parameter gather) -- they get
// the byte budget of a launch first.  [1] background: jobs with a far deadline (gradient reduce-scatters, the lm_head
// gather that is only needed after the last layer) -- they use whatever budget the foreground leaves.
std::deque<Job> g_queues[2];
long long g_next_id = 1;
long long g_stats[4] = {0, 0, 0, 0};
unsigned long long* g_debug = nullptr;   // device buffer of g_debug_cap records (8 x u64 each), see CarryArgs::debug
long long g_debug_cap = 0, g_debug_next = 0;
double g_bpf = -1.0;

double bpf_locked() {
  if (g_bpf < 0) {
    const char* e = getenv("TORCHACC_B200_CARRY_BYTES_PER_FLOP");
    g_bpf = e ? atof(e) : 1.2e-4;
    if (!(g_bpf >= 0)) g_bpf = 1.2e-4;
  }
  return g_bpf;
}

// hand out chunks [next, next + n) of the head job as one slice
void fill_slice(Job& j, uint32_t n, CarrySlice& out) {
  out = j.proto;
  out.chunk_begin = j.next;
  out.chunk_end = j.next + n;
  out.signal_entry = 0;      // published by symm_signal when the source data became final
  out.signal_exit = (j.next + n == j.total) ? 1 : 0;
  j.next += n;
}

}  // namespace

double carry_bytes_per_flop(double v) {
  std::lock_guard<std::mutex> lk(g_mu);
  if (v >= 0) g_bpf = v;
  return bpf_locked();
}

struct SignalArgs {
  uint32_t* pads[kCarryMaxWorld];
};
__global__ void __launch_bounds__(32) signal_entry_kernel(const SignalArgs a, int rank, int world, int channel,
                                                          uint32_t epoch) {
  if ((int)threadIdx.x < world) {
    __threadfence_system();
    carry_st_release_sys(a.pads[threadIdx.x] + channel * 16 + rank, epoch);
  }
}

cudaError_t symm_signal(const uint64_t* pad_ptrs, int rank, int world, int channel, uint32_t epoch, cudaStream_t stream) {
  if (world > kCarryMaxWorld || rank < 0 || rank >= world) return cudaErrorInvalidValue;
  SignalArgs a;
  for (int r = 0; r < kCarryMaxWorld; ++r) a.pads[r] = r < world ? reinterpret_cast<uint32_t*>(pad_ptrs[r]) : nullptr;
  signal_entry_kernel<<<1, 32, 0, stream>>>(a, rank, world, channel, epoch);
  return cudaGetLastError();
}

long long carry_push(int kind, const uint64_t* src, uint64_t dst, const uint64_t* pads, long long bytes, int rank,
                     int world, int channel, uint32_t epoch, uint64_t block_counter, float scale, int in_bf16,
                     int out_fp32, int accumulate, uint64_t stats, int background, int entry_channel,
                     uint32_t entry_epoch) {
  if ((kind != 1 && kind != 2) || world < 2 || world > kCarryMaxWorld || bytes <= 0 || bytes % 16 != 0) return -1;
  Job j;
  memset(&j.proto, 0, sizeof(j.proto));
  CarrySlice& p = j.proto;
  p.kind = kind; p.rank = rank; p.world = world; p.channel = channel; p.epoch = epoch;
  p.entry_channel = entry_channel; p.entry_epoch = entry_epoch;
  for (int r = 0; r < world; ++r) {
    p.src[r] = reinterpret_cast<const uint8_t*>(src[r]) + (kind == 2 ? (long long)rank * bytes : 0ll);
    p.pads[r] = reinterpret_cast<uint32_t*>(pads[r]);
  }
  p.dst = reinterpret_cast<uint8_t*>(dst);
  p.bytes = bytes;
  p.block_counter = reinterpret_cast<uint32_t*>(block_counter);
  p.scale = scale; p.in_bf16 = in_bf16; p.out_fp32 = out_fp32; p.accumulate = accumulate;
  p.stats = reinterpret_cast<float*>(stats);
  if (kind == 1) {
    p.chunk_bytes = kCarryStageBytes;
    const long long cps = (bytes + p.chunk_bytes - 1) / p.chunk_bytes;
    j.total = (uint32_t)(cps * world);              // own shard included: no separate local copy
    j.bytes_per_chunk = (double)p.chunk_bytes * (world - 1) / world;   // the own shard's chunks never leave the GPU
  } else {
    p.chunk_bytes = (kCarryStageBytes / (uint32_t)world) & ~127u;        // one ring stage holds `world` sub-chunks
    j.total = (uint32_t)((bytes + p.chunk_bytes - 1) / p.chunk_bytes);
    j.bytes_per_chunk = (double)p.chunk_bytes * world;      // every source passes through the reducing warp
  }
  j.next = 0;
  std::lock_guard<std::mutex> lk(g_mu);
  j.id = g_next_id++;
  g_queues[background ? 1 : 0].push_back(j);
  return j.id;
}

int carry_take(double flops, CarryArgs* out) {
  std::lock_guard<std::mutex> lk(g_mu);
  memset(out, 0, sizeof(*out));
  if (g_queues[0].empty() && g_queues[1].empty()) return 0;
  double budget = flops * bpf_locked();
  int n = 0;
  for (int q = 0; q < 2 && n < kCarrySlots; ++q) {
    std::deque<Job>& Q = g_queues[q];
    while (n < kCarrySlots && !Q.empty()) {
      Job& j = Q.front();
      const uint32_t left = j.total - j.next;
      uint32_t take = (uint32_t)(budget / j.bytes_per_chunk);
      if (take > left) take = left;
      if (take == 0) break;
      fill_slice(j, take, out->slice[n++]);
      g_stats[0] += take;
      budget -= take * j.bytes_per_chunk;
      if (j.next == j.total) Q.pop_front();
      else break;
    }
  }
  if (n > 0) {
    ++g_stats[2];
    if (g_debug != nullptr && g_debug_next < g_debug_cap) out->debug = g_debug + 8 * g_debug_next++;
  }
  return n;
}

int carry_take_probe(double flops, long long* out) {
  CarryArgs ca;
  const int n = carry_take(flops, &ca);
  for (int i = 0; i < kCarrySlots; ++i) {
    out[4 * i + 0] = i < n ? ca.slice[i].kind : 0;
    out[4 * i + 1] = i < n ? ca.slice[i].chunk_begin : 0;
    out[4 * i + 2] = i < n ? ca.slice[i].chunk_end : 0;
    out[4 * i + 3] = i < n ? ca.slice[i].chunk_bytes : 0;
  }
  return n;
}

long long carry_pending(long long job_id, int queue) {
  std::lock_guard<std::mutex> lk(g_mu);
  long long c = 0;
  for (int q = 0; q < 2; ++q) {
    if (queue >= 0 && q != queue) continue;
    for (const Job& j : g_queues[q])
      if (job_id == 0 || j.id <= job_id) c += j.total - j.next;
  }
  return c;
}

long long carry_set_debug(unsigned long long* buf, long long records) {
  std::lock_guard<std::mutex> lk(g_mu);
  const long long used = g_debug_next;
  g_debug = buf; g_debug_cap = records; g_debug_next = 0;
  return used;
}

void carry_stats(long long* out4, int reset) {
  std::lock_guard<std::mutex> lk(g_mu);
  for (int i = 0; i < 4; ++i) out4[i] = g_stats[i];
  if (reset) memset(g_stats, 0, sizeof(g_stats));
}

// ------------------------------------------------------------------------------------------------------
// stand-alone execution of whatever no GEMM could carry (first gather of a step, last reductions of a backward)
// ------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(32)
carry_only_kernel(const CarryArgs ca) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t ring = (smem_u32(smem_raw) + 127u) & ~127u;
  carry_role(ca, ring, ring + kCarryStages * kCarryStageBytes, (int)blockIdx.x, (int)gridDim.x);
}

cudaError_t carry_flush(long long job_id, int queue, int num_sms, cudaStream_t stream) {
  static cudaError_t cfg = cudaFuncSetAttribute(carry_only_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                kCarrySmemBytes + 128);
  if (cfg != cudaSuccess) return cfg;
  static int ctas_per_sm = [] {
    const char* e = getenv("TORCHACC_B200_CARRY_FLUSH_CTAS_PER_SM");
    const int v = e ? atoi(e) : 2;
    return v > 0 ? v : 2;
  }();
  for (;;) {
    CarryArgs ca;
    memset(&ca, 0, sizeof(ca));
    int n = 0;
    {
      std::lock_guard<std::mutex> lk(g_mu);
      for (int q = 0; q < 2; ++q) {
        if (queue >= 0 && q != queue) continue;
        std::deque<Job>& Q = g_queues[q];
        while (n < kCarrySlots && !Q.empty()) {
          Job& j = Q.front();
          if (job_id != 0 && j.id > job_id) break;
          const uint32_t left = j.total - j.next;
          fill_slice(j, left, ca.slice[n++]);
          g_stats[1] += left;
          Q.pop_front();
        }
      }
      if (n > 0) ++g_stats[3];
    }
    if (n == 0) return cudaSuccess;
    carry_only_kernel<<<num_sms * ctas_per_sm, 32, kCarrySmemBytes + 128, stream>>>(ca);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
  }
}

__global__ void __launch_bounds__(32)
wait_done_kernel(const uint32_t* pad, int rank, int world, int channel, uint32_t epoch) {
  if ((int)threadIdx.x < world)
    spin_until_epoch(pad + channel * 16 + 8 + threadIdx.x, epoch, rank, (int)threadIdx.x, channel, "carry done");
}

cudaError_t symm_wait_done(const uint64_t* pad_ptrs, int rank, int world, int channel, uint32_t epoch,
                           cudaStream_t stream) {
  if (world > kCarryMaxWorld || rank < 0 || rank >= world) return cudaErrorInvalidValue;
  wait_done_kernel<<<1, 32, 0, stream>>>(reinterpret_cast<const uint32_t*>(pad_ptrs[rank]), rank, world, channel, epoch);
  return cudaGetLastError();
}

}  // namespace tb
