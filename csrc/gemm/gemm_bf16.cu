// Persistent, warp-specialised tcgen05 GEMM for sm_100a.
//
//   D[M,N] (row-major) (+)= A_op[M,K] * B_op[N,K]^T (+ bias[N])
//
// Each operand may be stored K-major (reduction dim contiguous) or MN-major (M / N contiguous), so the
// one kernel covers the three GEMMs of a linear layer without any transposes in memory:
//   forward  y  = x  W^T : A = x  (K-major),  B = W  (K-major)
//   dgrad    dx = dy W   : A = dy (K-major),  B = W  (MN-major)
//   wgrad    dW = dy^T x : A = dy (MN-major), B = x  (MN-major)
//
// Structure (one CTA per SM, or one CTA *pair* per 2 SMs with cta_group::2):
//   warp 0   : TMA producer   (cp.async.bulk.tensor -> 128B-swizzled smem ring, mbarrier full/empty)
//   warp 1   : MMA issuer     (one elected thread; tcgen05.mma kind::f16, fp32 accumulators in TMEM)
//   warp 2   : TMEM allocator (512 columns = 2 accumulator stages of 128 lanes x 256 columns)
//   warps 4-7: epilogue       (tcgen05.ld -> registers -> bias/accumulate/cast -> 16B global stores)
// The accumulator is double-buffered in TMEM so the epilogue of tile i overlaps the mainloop of tile i+1.
// With kCluster == 2 the pair computes a 256x256 tile: each CTA stages its own 128 rows of A and 128 of the
// 256 rows of B, halving shared-memory and L2 traffic per FLOP; only the leader CTA issues MMAs.
//
// Reference parity: replaces the cuBLAS GEMMs the reference reaches through nn.Linear
// (reference torchacc/__init__.py:97-98 forces XLA onto cuBLAS; eager path uses cuBLASLt).
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../common/ptx.cuh"
#include "../common/tensormap.h"
#include "../fused/carry.cuh"
#include "../fused/carry.h"
#include "gemm.h"

namespace tb {

constexpr int kBlockMCta = 128;  // accumulator rows per CTA (TMEM lanes)
constexpr int kBlockN = 256;     // UMMA N
constexpr int kBlockK = 64;      // 128 bytes of bf16 = one swizzle row
constexpr int kUmmaK = 16;
constexpr int kNumThreads = 256;
constexpr int kGroupM = 8;       // tile rasterisation: sweep 8 row-tiles before moving along N

constexpr int kMaxPeers = 8;

// Fusion of the GEMM with the adjacent tensor-parallel collective over NVLink peer memory.
//   kFuse == 1 (all-gather -> GEMM): the A operand is the token-gathered activation.  Every rank owns a symmetric
//     [world*rows, K] buffer and has already written ITS row block; the leading `comm_clusters` clusters of this
//     kernel pull the other ranks' row blocks from peer memory into the local buffer and publish one ready counter
//     per source rank; the GEMM clusters consume M-tiles in shard order (own block first) and their TMA producers
//     spin on the counter of the shard they are about to read.  The transfer of shard r+1 overlaps the MMAs of shard r.
//   kFuse == 2 (GEMM -> reduce-scatter): output row block d belongs to rank d.  The epilogue stores the bf16 partial
//     tile straight into rank d's staging slot [my_rank] with peer stores and bumps a per-(dst, src) arrival counter
//     (release.sys); a small reduce kernel on each rank sums the `world` slots in fp32 once all sources have arrived.
struct FuseArgs {
  void* peer[kMaxPeers];          // kFuse 1: peers' gathered-A buffers; kFuse 2: peers' staging buffers
  uint32_t* pads[kMaxPeers];      // symmetric signal pads (entry / exit barriers)
  uint32_t* peer_cnt[kMaxPeers];  // kFuse 2: peers' arrival counters (array of `world` uint32 per rank)
  void* a_full;                   // kFuse 1: local gathered buffer
  uint32_t* flags;                // kFuse 1: local per-source ready counters
  uint32_t* block_counter;        // exit barrier bookkeeping (local)
  int rank, world, rows_per_rank, channel;
  uint32_t epoch;                 // monotonically increasing per call on this channel
  uint32_t flag_target;           // kFuse 1: counter value that means "shard complete" for this call
  int comm_clusters;              // kFuse 1: leading clusters acting as copy engines
  long long slot_stride;          // kFuse 2: elements between two source slots of a staging buffer
};

struct GemmArgs {
  void* D;
  const __nv_bfloat16* bias;
  int M, N, K;
  long long ldd;
  const void* C;   // optional addend with D's dtype: D = result (+ bias) + C.  C == D is gradient accumulation;
  long long ldc;   // a different C fuses a residual add into the epilogue (h + x W^T).
  int out_fp32;
  int num_m_tiles, num_n_tiles;
  // SwiGLU epilogue (gate|up projection, cta_group::2 only): B is W_gu [2F][K]; CTA 0 of a pair stages 128 gate rows,
  // CTA 1 the matching 128 up rows, so accumulator columns [0,128) / [128,256) are gate / up of the SAME 128 output
  // columns and the epilogue writes h = silu(g) * u [M][F] directly (and g|u [M][2F] only when the backward needs it).
  int swiglu;      // 0 / 1
  int swiglu_F;    // F = intermediate size
  void* gu_out;    // optional [M][2F] pre-activation (nullptr: not materialised)
  long long ldgu;
  int fp16;        // operands, bias and 16-bit outputs are IEEE fp16 instead of bf16 (same tiles, other format bits)
  int dynamic;     // 1: grid = one cluster per tile, tiles are claimed with cluster-launch-control (see kernel comment)
  FuseArgs fuse;
  CarryArgs carry;  // slices of pending FSDP collectives moved by warp 3 of every CTA while the tiles run (fused/carry.cuh)
};

TB_DEVICE void st_release_sys_u32(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
TB_DEVICE uint32_t ld_acquire_sys_u32(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
TB_DEVICE uint32_t ld_acquire_gpu_u32(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
TB_DEVICE void red_release_sys_add(uint32_t* p, uint32_t v) {
  asm volatile("red.release.sys.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
TB_DEVICE void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }
TB_DEVICE uint4 ld_volatile_v4(const void* p) {
  uint4 v;
  asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];"
               : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
               : "l"(p)
               : "memory");
  return v;
}

// Cross-rank barrier on the symmetric signal pad (slot layout as in comm/symm_comm.cu: 16 words per channel,
// [0,8) entry flags, [8,16) exit flags).  `signal` is true for exactly one CTA per rank.
// `tid` enumerates the calling threads (threadIdx.x for a whole CTA, the lane id for a single warp).
TB_DEVICE void fused_barrier(const FuseArgs& f, bool signal, int slot_base, int tid) {
  if (signal && tid < f.world) {
    __threadfence_system();
    st_release_sys_u32(f.pads[tid] + f.channel * 16 + slot_base + f.rank, f.epoch);
  }
  if (tid < f.world)
    spin_until_epoch(f.pads[f.rank] + f.channel * 16 + slot_base + tid, f.epoch, f.rank, tid, f.channel,
                     "fused GEMM barrier");
}

// Copy-engine role of the all-gather -> GEMM kernel (executed by whole clusters).  One thread per CTA drives a ring
// of 32 KB TMA bulk copies  peer HBM --NVLink--> smem --> local gathered buffer:  (stages-1) x 32 KB are in flight per
// CTA, so a handful of CTAs keep the link full (16-byte register loads would need ~100 CTAs for the same bytes in
// flight).  Shard r is announced to the MMA clusters' TMA producers through flags[r] when all copy CTAs finished it.
TB_DEVICE void gather_role(const FuseArgs& f, int K, int comm_cta, int num_comm_ctas, uint32_t smem_base,
                           uint32_t bar_base, int num_stages) {
  constexpr uint32_t kChunk = 32768;
  if (threadIdx.x == 0) {
    for (int s = 0; s < num_stages; ++s) mbar_init(bar_base + 8u * s, 1);
    fence_mbar_init();
  }
  fused_barrier(f, comm_cta == 0, 0, threadIdx.x);   // every rank has written its own row block
  __syncthreads();
  if (threadIdx.x == 0 && f.world > 1) {
    const size_t shard_bytes = (size_t)f.rows_per_rank * K * 2;
    const int chunks = (int)(shard_bytes / kChunk);            // rows % 256 == 0 and K % 64 == 0 -> exact
    const int per = chunks > comm_cta ? (chunks - comm_cta + num_comm_ctas - 1) / num_comm_ctas : 0;
    const int total = per * (f.world - 1);
    auto locate = [&](int j, const uint8_t*& src, uint8_t*& dst) {
      const int r = (f.rank + 1 + j / per) % f.world;
      const size_t off = (size_t)r * shard_bytes + (size_t)(comm_cta + (j % per) * num_comm_ctas) * kChunk;
      src = reinterpret_cast<const uint8_t*>(f.peer[r]) + off;
      dst = reinterpret_cast<uint8_t*>(f.a_full) + off;
    };
    auto issue = [&](int j) {
      const uint8_t* src; uint8_t* dst;
      locate(j, src, dst);
      const int s = j % num_stages;
      mbar_arrive_expect_tx(bar_base + 8u * s, kChunk);
      bulk_load(smem_base + s * kChunk, src, kChunk, bar_base + 8u * s);
    };
    int issued = 0;
    for (; issued < num_stages - 1 && issued < total; ++issued) issue(issued);
    for (int i = 0; i < total; ++i) {
      const int s = i % num_stages;
      mbar_wait(bar_base + 8u * s, (i / num_stages) & 1);
      const uint8_t* src; uint8_t* dst;
      locate(i, src, dst);
      bulk_store(dst, smem_base + s * kChunk, kChunk);
      tma_store_commit();
      if (issued < total) {
        tma_store_wait_read<1>();       // every store but the newest has left smem -> stage (i-1) % stages is free
        issue(issued++);
      }
      if ((i + 1) % per == 0) {         // last chunk of a source rank handled by this CTA
        tma_store_wait<0>();            // writes complete
        __threadfence();
        atomicAdd(f.flags + (f.rank + 1 + i / per) % f.world, 1u);
      }
    }
    if (per == 0)                       // more copy CTAs than chunks: still count towards every shard's flag
      for (int st = 1; st < f.world; ++st) atomicAdd(f.flags + (f.rank + st) % f.world, 1u);
  }
  // exit: all my copy CTAs are done reading peers -> tell them; stay until every peer is done reading me
  __shared__ uint32_t s_last;
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    s_last = (atomicAdd(f.block_counter, 1u) == (uint32_t)num_comm_ctas - 1) ? 1u : 0u;
  }
  __syncthreads();
  if (!s_last) return;
  if (threadIdx.x == 0) *f.block_counter = 0;
  fused_barrier(f, true, 8, threadIdx.x);
}

template <int kCluster>
struct GemmSmem {
  static constexpr int kLoadN = kBlockN / kCluster;
  static constexpr int kABytes = kBlockMCta * kBlockK * 2;
  static constexpr int kBBytes = kLoadN * kBlockK * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kStages = (kCluster == 2) ? 6 : 4;
  static constexpr int kBarrierBytes = 1024;
  static constexpr int kTotal = kStages * kStageBytes + kBarrierBytes + 1024;  // +1024 alignment slack
  static constexpr int kTotalCarry = kTotal + kCarrySmemBytes;                 // + the copy warp's private ring
};

__device__ __forceinline__ void tile_coords(int t, int num_m_tiles, int num_n_tiles, int& tm, int& tn) {
  const int per_group = kGroupM * num_n_tiles;
  const int g = t / per_group;
  const int first_m = g * kGroupM;
  const int gsize = min(kGroupM, num_m_tiles - first_m);
  const int r = t - g * per_group;
  tm = first_m + (r % gsize);
  tn = r / gsize;
}

// tile order for the fused variants: row blocks ("shards") are visited in ring order starting from `first`
__device__ __forceinline__ void tile_coords_sharded(int t, int tiles_m_per_shard, int num_n_tiles, int first, int world,
                                                    int& tm, int& tn, int& shard) {
  const int per_shard = tiles_m_per_shard * num_n_tiles;
  const int si = t / per_shard;
  shard = (first + si) % world;
  int lm, ln;
  tile_coords(t - si * per_shard, tiles_m_per_shard, num_n_tiles, lm, ln);
  tm = shard * tiles_m_per_shard + lm;
  tn = ln;
}

template <int kCluster, Major kAMajor, Major kBMajor, int kFuse>
__global__ void __launch_bounds__(kNumThreads, 1)
gemm_bf16_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                 const GemmArgs args) {
  using S = GemmSmem<kCluster>;
  constexpr int kStages = S::kStages;
  constexpr int kLoadN = S::kLoadN;
  constexpr uint32_t kUmmaM = kBlockMCta * kCluster;
  constexpr uint32_t kIdesc = make_idesc_f16(kUmmaM, kBlockN, kAMajor, kBMajor, true);

  extern __shared__ uint8_t smem_raw[];
  // SWIZZLE_128B tiles need 1024-byte alignment.
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bar_base = smem_base + kStages * S::kStageBytes;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (kStages + s); };
  auto tfull_bar = [&](int s) { return bar_base + 8u * (2 * kStages + s); };
  auto tempty_bar = [&](int s) { return bar_base + 8u * (2 * kStages + 2 + s); };
  const uint32_t tmem_slot = bar_base + 8u * (2 * kStages + 4);
  // dynamic tile scheduler (cluster launch control): 16-byte responses + full/empty barriers
  constexpr int kSched = 4;
  auto clc_full = [&](int s) { return bar_base + 8u * (2 * kStages + 6 + s); };
  auto clc_empty = [&](int s) { return bar_base + 8u * (2 * kStages + 6 + kSched + s); };
  auto clc_resp = [&](int s) { return bar_base + 512u + 16u * s; };
  constexpr uint32_t kSchedConsumers = 5 * kCluster + 1;   // per CTA: TMA warp + 4 epilogue warps; + the MMA warp
  auto smem_a = [&](int s) { return smem_base + s * S::kStageBytes; };
  auto smem_b = [&](int s) { return smem_base + s * S::kStageBytes + S::kABytes; };

  const int warp_idx = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  const uint32_t lane = lane_id();
  const uint32_t cta_rank = (kCluster == 2) ? cluster_ctarank() : 0u;
  const bool is_leader = cta_rank == 0;

  const int num_tiles = args.num_m_tiles * args.num_n_tiles;
  const int num_k_blocks = (args.K + kBlockK - 1) / kBlockK;
  int cluster_id = blockIdx.x / kCluster;
  int num_clusters = gridDim.x / kCluster;
  if constexpr (kFuse == 1) {
    const int cc = args.fuse.comm_clusters;
    if (cluster_id < cc) {   // copy-engine clusters: no TMEM, no MMA
      gather_role(args.fuse, args.K, (int)blockIdx.x, cc * kCluster, smem_base, bar_base,
                  kStages * S::kStageBytes / 32768);
      return;
    }
    cluster_id -= cc;
    num_clusters -= cc;
  }
  const int tiles_m_per_shard = (kFuse != 0) ? args.fuse.rows_per_rank / (kBlockMCta * kCluster) : 0;
  // fused tile order: all-gather starts with the local row block; reduce-scatter ends with it (peers first)
  const int first_shard = (kFuse == 1) ? args.fuse.rank : ((kFuse == 2) ? (args.fuse.rank + 1) % args.fuse.world : 0);
  auto coords = [&](int t, int& tm, int& tn, int& shard) {
    if constexpr (kFuse != 0) {
      tile_coords_sharded(t, tiles_m_per_shard, args.num_n_tiles, first_shard, args.fuse.world, tm, tn, shard);
    } else {
      shard = 0;
      tile_coords(t, args.num_m_tiles, args.num_n_tiles, tm, tn);
    }
  };

  if (warp_idx == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_b);
  }
  if (warp_idx == 1 && lane == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(tfull_bar(s), 1);
      mbar_init(tempty_bar(s), 4 * kCluster);  // one arrive per epilogue warp (of both CTAs)
    }
    for (int s = 0; s < kSched; ++s) {
      mbar_init(clc_full(s), 1);
      mbar_init(clc_empty(s), kSchedConsumers);
    }
    fence_mbar_init();
  }
  if (warp_idx == 2) tmem_alloc<kCluster>(tmem_slot, 512);
  tc_fence_before();
  if constexpr (kCluster == 2) cluster_sync(); else __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

  // ---- tile iteration ----
  // static : tile = cluster_id, += num_clusters (one resident cluster per SM pair for the whole kernel)
  // dynamic: the grid has one cluster per tile.  A launched cluster starts with the tile of its own index and then
  //          claims not-yet-launched clusters through cluster launch control until none is left.  SMs that are busy
  //          with another kernel (a collective on a side stream) simply join late or never: the remaining clusters
  //          absorb their tiles instead of the whole GEMM waiting for stragglers.
  const bool dynamic = (kFuse == 0) && args.dynamic != 0;
  int sched_it = 0;   // per role thread: position in the response pipeline
  auto advance = [&](int& tile, bool do_arrive) -> bool {
    if (!dynamic) {
      tile += num_clusters;
      return tile < num_tiles;
    }
    const int s = sched_it % kSched;
    mbar_wait(clc_full(s), (sched_it / kSched) & 1);
    uint32_t x;
    const bool valid = clc_query(clc_resp(s), x);
    fence_proxy_async_smem();   // our read of the response precedes the next asynchronous write of this slot
    if (do_arrive) {
      if constexpr (kCluster == 2) mbar_arrive_cluster(mapa_shared(clc_empty(s), 0));
      else mbar_arrive(clc_empty(s));
    }
    ++sched_it;
    tile = (int)(x / kCluster);
    return valid;
  };
  if (dynamic && warp_idx == 3 && lane == 0) {
    // ================================ Tile scheduler ================================
    for (int it = 0;; ++it) {
      const int s = it % kSched;
      if (is_leader && it >= kSched) {   // every consumer (of both CTAs) has read the previous response in this slot
        if constexpr (kCluster == 2) mbar_wait_cluster(clc_empty(s), ((it / kSched) - 1) & 1);
        else mbar_wait(clc_empty(s), ((it / kSched) - 1) & 1);
      }
      mbar_arrive_expect_tx(clc_full(s), 16);       // each CTA arms its own barrier; the response is multicast
      if (is_leader) clc_try_cancel<kCluster == 2>(clc_resp(s), clc_full(s));
      mbar_wait(clc_full(s), (it / kSched) & 1);
      uint32_t x;
      if (!clc_query(clc_resp(s), x)) break;
    }
  }

  if (warp_idx == 0) {
    // ================================ TMA producer ================================
    if (lane == 0) {
      int it = 0;
      int t = cluster_id;
      for (bool more = t < num_tiles; more; more = advance(t, true)) {
        int tm, tn, shard;
        coords(t, tm, tn, shard);
        const int m0 = tm * (int)kUmmaM + (int)cta_rank * kBlockMCta;
        const int n0 = (kFuse == 0 && args.swiglu) ? tn * kLoadN + (int)cta_rank * args.swiglu_F
                                                   : tn * kBlockN + (int)cta_rank * kLoadN;
        if constexpr (kFuse == 1) {
          if (shard != args.fuse.rank) {   // wait until the copy clusters have landed this source rank's rows
            spin_until_count(args.fuse.flags + shard, args.fuse.flag_target, args.fuse.rank, shard,
                             "all-gather -> GEMM shard flag");
            fence_proxy_async_all();       // generic-proxy writes of the copy CTAs -> visible to TMA
          }
        }
        for (int kb = 0; kb < num_k_blocks; ++kb, ++it) {
          const int s = it % kStages;
          const uint32_t ph = (it / kStages) & 1;
          mbar_wait(empty_bar(s), ph ^ 1);
          if (is_leader) mbar_arrive_expect_tx(full_bar(s), S::kStageBytes * kCluster);
          const int k0 = kb * kBlockK;
          auto load = [&](uint32_t dst, const CUtensorMap* map, int c0, int c1) {
            if constexpr (kCluster == 2) tma_load_2d_2sm(dst, map, full_bar(s), c0, c1);
            else tma_load_2d(dst, map, full_bar(s), c0, c1);
          };
          if constexpr (kAMajor == Major::K) {
            load(smem_a(s), &tmap_a, k0, m0);
          } else {
#pragma unroll
            for (int i = 0; i < kBlockMCta / 64; ++i) load(smem_a(s) + i * 8192, &tmap_a, m0 + i * 64, k0);
          }
          if constexpr (kBMajor == Major::K) {
            load(smem_b(s), &tmap_b, k0, n0);
          } else {
#pragma unroll
            for (int i = 0; i < kLoadN / 64; ++i) load(smem_b(s) + i * 8192, &tmap_b, n0 + i * 64, k0);
          }
        }
      }
    }
  } else if (warp_idx == 1) {
    // ================================ MMA issuer ================================
    if (is_leader && lane == 0) {
      // a/b format bits [7,10) / [10,13): 1 = bf16, 0 = fp16
      const uint32_t idesc = args.fp16 ? (kIdesc & ~((7u << 7) | (7u << 10))) : kIdesc;
      int it = 0, lt = 0;
      int t = cluster_id;
      for (bool more = t < num_tiles; more; more = advance(t, true), ++lt) {
        const int as = lt & 1;
        const uint32_t aph = (lt >> 1) & 1;
        if constexpr (kCluster == 2) mbar_wait_cluster(tempty_bar(as), aph ^ 1);
        else mbar_wait(tempty_bar(as), aph ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + as * kBlockN;
        for (int kb = 0; kb < num_k_blocks; ++kb, ++it) {
          const int s = it % kStages;
          const uint32_t ph = (it / kStages) & 1;
          mbar_wait(full_bar(s), ph);
          tc_fence_after();
#pragma unroll
          for (int k = 0; k < kBlockK / kUmmaK; ++k) {
            const uint64_t da = (kAMajor == Major::K) ? desc_kmajor_sw128(smem_a(s), k)
                                                       : desc_mnmajor_sw128(smem_a(s), k, 8192);
            const uint64_t db = (kBMajor == Major::K) ? desc_kmajor_sw128(smem_b(s), k)
                                                       : desc_mnmajor_sw128(smem_b(s), k, 8192);
            umma_ss_f16<kCluster>(tmem_d, da, db, idesc, (kb | k) != 0);
          }
          // Release the smem slot once the MMAs that read it retire.
          if constexpr (kCluster == 2) umma_commit_2sm(empty_bar(s), 0x3);
          else umma_commit(empty_bar(s));
        }
        if constexpr (kCluster == 2) umma_commit_2sm(tfull_bar(as), 0x3);
        else umma_commit(tfull_bar(as));
      }
    }
  } else if (warp_idx == 3) {
    // ================================ Carried collective ================================
    // Under static scheduling this warp has no GEMM role: it moves this launch's slice of the pending FSDP
    // all-gather / reduce-scatter (TMA bulk copies through a private 24 KB ring) while the tiles are computed.
    if constexpr (kFuse == 0) {
      if (!dynamic && args.carry.slice[0].kind != 0) {
        const uint32_t ring = bar_base + S::kBarrierBytes;
        carry_role(args.carry, ring, ring + kCarryStages * kCarryStageBytes, (int)blockIdx.x, (int)gridDim.x);
      }
    }
  } else if (warp_idx >= 4) {
    // ================================ Epilogue ================================
    const uint32_t q = warp_idx & 3;  // TMEM lane quarter this warp may read
    const bool f16 = args.fp16 != 0;
    auto unpk = [&](uint32_t u) { return f16 ? unpack_f16x2(u) : unpack_bf16x2(u); };
    auto pk = [&](float a, float b) { return f16 ? pack_f16x2(a, b) : pack_bf16x2(a, b); };
    auto ld1 = [&](const __nv_bfloat16* p) {
      return f16 ? __half2float(*reinterpret_cast<const __half*>(p)) : __bfloat162float(*p);
    };
    auto st1 = [&](__nv_bfloat16* p, float v) {
      if (f16) *reinterpret_cast<__half*>(p) = __float2half_rn(v); else *p = __float2bfloat16(v);
    };
    const uint32_t tempty_leader =
        (kCluster == 2) ? mapa_shared(tempty_bar(0), 0) : tempty_bar(0);  // barrier lives in the leader CTA
    int lt = 0;
    if constexpr (kFuse == 2) {
      // nobody may write into a peer's staging buffer before that peer finished reducing the previous call
      fused_barrier(args.fuse, blockIdx.x == 0 && warp_idx == 4, 0, (int)lane);
      __syncwarp();
    }
    // kFuse == 2: arrivals are batched per destination rank -- ONE system-scope fence + counter add per warp and
    // shard instead of one per tile (a fence.sys costs an NVLink round trip, longer than a whole K=2048 mainloop).
    int cur_shard = -1;
    uint32_t pending = 0;
    auto flush_arrivals = [&]() {
      if constexpr (kFuse == 2) {
        if (pending != 0) {
          __threadfence_system();
          __syncwarp();
          if (lane == 0) red_release_sys_add(args.fuse.peer_cnt[cur_shard] + args.fuse.rank, pending);
          pending = 0;
        }
      }
    };
    int t = cluster_id;
    for (bool more = t < num_tiles; more; more = advance(t, lane == 0), ++lt) {
      int tm, tn, shard;
      coords(t, tm, tn, shard);
      const int as = lt & 1;
      const uint32_t aph = (lt >> 1) & 1;
      mbar_wait(tfull_bar(as), aph);
      tc_fence_after();
      const long long grow = (long long)tm * kUmmaM + cta_rank * kBlockMCta + q * 32 + lane;
      const int n0 = tn * kBlockN;
      const bool row_ok = grow < args.M;
      if constexpr (kFuse == 2) {
        // bf16 partial tile -> staging slot [my rank] of the rank that owns this row block (peer stores over NVLink).
        // The staging layout is block-major: one 2 KB block per (tile, CTA half, lane quarter, 32-column chunk),
        // inside a block [j][lane][8 columns], so every warp-wide store instruction writes 512 contiguous bytes
        // (full NVLink packets) instead of 32 scattered 16-byte pieces.  rs_reduce_kernel undoes the permutation.
        if (shard != cur_shard) {
          flush_arrivals();
          cur_shard = shard;
        }
        const int tm_local = tm - shard * tiles_m_per_shard;
        const long long blk0 = ((((long long)tm_local * args.num_n_tiles + tn) * 2 + cta_rank) * 4 + q) * 8;
        uint4* dbase = reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(args.fuse.peer[shard]) +
                                                (long long)args.fuse.rank * args.fuse.slot_stride) + blk0 * 128 + lane;
#pragma unroll 1
        for (int c = 0; c < kBlockN / 32; ++c) {
          __syncwarp();
          uint32_t r[32];
          tmem_ld_32x32b_x32(tmem_base + ((q * 32u) << 16) + as * kBlockN + c * 32, r);
          tmem_ld_wait();
          if (n0 + c * 32 + 32 > args.N) continue;     // (rows are always in range: rows_per_rank % 256 == 0)
          uint4* d4 = dbase + c * 128;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            uint4 o;
            o.x = pack_bf16x2(__uint_as_float(r[8 * j + 0]), __uint_as_float(r[8 * j + 1]));
            o.y = pack_bf16x2(__uint_as_float(r[8 * j + 2]), __uint_as_float(r[8 * j + 3]));
            o.z = pack_bf16x2(__uint_as_float(r[8 * j + 4]), __uint_as_float(r[8 * j + 5]));
            o.w = pack_bf16x2(__uint_as_float(r[8 * j + 6]), __uint_as_float(r[8 * j + 7]));
            d4[j * 32] = o;
          }
        }
        ++pending;
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          if constexpr (kCluster == 2) mbar_arrive_cluster(tempty_leader + 8u * as);
          else mbar_arrive(tempty_leader + 8u * as);
        }
        continue;
      }
      if constexpr (kFuse == 0 && kCluster == 2) {
        if (args.swiglu) {
          const int oc0 = tn * 128;     // output columns of this tile
#pragma unroll 1
          for (int c = 0; c < 4; ++c) {
            __syncwarp();
            uint32_t rg[32], ru[32];
            tmem_ld_32x32b_x32(tmem_base + ((q * 32u) << 16) + as * kBlockN + c * 32, rg);
            tmem_ld_32x32b_x32(tmem_base + ((q * 32u) << 16) + as * kBlockN + 128 + c * 32, ru);
            tmem_ld_wait();
            const int gcol = oc0 + c * 32;
            if (!row_ok || gcol >= args.swiglu_F) continue;
            __nv_bfloat16* hp = reinterpret_cast<__nv_bfloat16*>(args.D) + grow * args.ldd + gcol;
            uint4* h4 = reinterpret_cast<uint4*>(hp);
            float hv[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              const float g = __uint_as_float(rg[j]), u = __uint_as_float(ru[j]);
              float t;
              asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(0.5f * g));     // sigmoid(g) = 0.5 + 0.5 tanh(g / 2)
              hv[j] = g * fmaf(0.5f, t, 0.5f) * u;
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              uint4 o;
              o.x = pk(hv[8 * j + 0], hv[8 * j + 1]); o.y = pk(hv[8 * j + 2], hv[8 * j + 3]);
              o.z = pk(hv[8 * j + 4], hv[8 * j + 5]); o.w = pk(hv[8 * j + 6], hv[8 * j + 7]);
              h4[j] = o;
            }
            if (args.gu_out != nullptr) {
              __nv_bfloat16* gp = reinterpret_cast<__nv_bfloat16*>(args.gu_out) + grow * args.ldgu + gcol;
              uint4* g4 = reinterpret_cast<uint4*>(gp);
              uint4* u4 = reinterpret_cast<uint4*>(gp + args.swiglu_F);
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                uint4 a, b;
                a.x = pk(__uint_as_float(rg[8 * j + 0]), __uint_as_float(rg[8 * j + 1]));
                a.y = pk(__uint_as_float(rg[8 * j + 2]), __uint_as_float(rg[8 * j + 3]));
                a.z = pk(__uint_as_float(rg[8 * j + 4]), __uint_as_float(rg[8 * j + 5]));
                a.w = pk(__uint_as_float(rg[8 * j + 6]), __uint_as_float(rg[8 * j + 7]));
                b.x = pk(__uint_as_float(ru[8 * j + 0]), __uint_as_float(ru[8 * j + 1]));
                b.y = pk(__uint_as_float(ru[8 * j + 2]), __uint_as_float(ru[8 * j + 3]));
                b.z = pk(__uint_as_float(ru[8 * j + 4]), __uint_as_float(ru[8 * j + 5]));
                b.w = pk(__uint_as_float(ru[8 * j + 6]), __uint_as_float(ru[8 * j + 7]));
                g4[j] = a;
                u4[j] = b;
              }
            }
          }
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive_cluster(tempty_leader + 8u * as);
          continue;
        }
      }
#pragma unroll 1
      for (int c = 0; c < kBlockN / 32; ++c) {
        __syncwarp();  // tcgen05.ld is warp-collective (.sync.aligned): reconverge after the guarded stores
        uint32_t r[32];
        tmem_ld_32x32b_x32(tmem_base + ((q * 32u) << 16) + as * kBlockN + c * 32, r);
        tmem_ld_wait();
        const int gcol = n0 + c * 32;
        if (!row_ok || gcol >= args.N) continue;
        float v[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
        const bool full = gcol + 32 <= args.N;
        if (args.bias != nullptr) {
          if (full) {
            const uint4* bp = reinterpret_cast<const uint4*>(args.bias + gcol);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              uint4 b = __ldg(bp + j);
              float2 f0 = unpk(b.x), f1 = unpk(b.y), f2 = unpk(b.z), f3 = unpk(b.w);
              v[8 * j + 0] += f0.x; v[8 * j + 1] += f0.y; v[8 * j + 2] += f1.x; v[8 * j + 3] += f1.y;
              v[8 * j + 4] += f2.x; v[8 * j + 5] += f2.y; v[8 * j + 6] += f3.x; v[8 * j + 7] += f3.y;
            }
          } else {
            for (int j = 0; j < 32; ++j)
              if (gcol + j < args.N) v[j] += ld1(args.bias + gcol + j);
          }
        }
        if (args.out_fp32) {
          float* dp = reinterpret_cast<float*>(args.D) + grow * args.ldd + gcol;
          const float* cp = args.C ? reinterpret_cast<const float*>(args.C) + grow * args.ldc + gcol : nullptr;
          if (full) {
            float4* d4 = reinterpret_cast<float4*>(dp);
            const float4* c4 = reinterpret_cast<const float4*>(cp);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              float4 o = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
              if (cp) {
                float4 old = c4[j];
                o.x += old.x; o.y += old.y; o.z += old.z; o.w += old.w;
              }
              d4[j] = o;
            }
          } else {
            for (int j = 0; j < 32; ++j)
              if (gcol + j < args.N) dp[j] = v[j] + (cp ? cp[j] : 0.f);
          }
        } else {
          __nv_bfloat16* dp = reinterpret_cast<__nv_bfloat16*>(args.D) + grow * args.ldd + gcol;
          const __nv_bfloat16* cp =
              args.C ? reinterpret_cast<const __nv_bfloat16*>(args.C) + grow * args.ldc + gcol : nullptr;
          if (full) {
            uint4* d4 = reinterpret_cast<uint4*>(dp);
            const uint4* c4 = reinterpret_cast<const uint4*>(cp);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              if (cp) {
                uint4 old = c4[j];
                float2 f0 = unpk(old.x), f1 = unpk(old.y), f2 = unpk(old.z), f3 = unpk(old.w);
                v[8 * j + 0] += f0.x; v[8 * j + 1] += f0.y; v[8 * j + 2] += f1.x; v[8 * j + 3] += f1.y;
                v[8 * j + 4] += f2.x; v[8 * j + 5] += f2.y; v[8 * j + 6] += f3.x; v[8 * j + 7] += f3.y;
              }
              uint4 o;
              o.x = pk(v[8 * j + 0], v[8 * j + 1]);
              o.y = pk(v[8 * j + 2], v[8 * j + 3]);
              o.z = pk(v[8 * j + 4], v[8 * j + 5]);
              o.w = pk(v[8 * j + 6], v[8 * j + 7]);
              d4[j] = o;
            }
          } else {
            for (int j = 0; j < 32; ++j)
              if (gcol + j < args.N) st1(dp + j, v[j] + (cp ? ld1(cp + j) : 0.f));
          }
        }
      }
      // All TMEM reads of this accumulator stage are complete: hand it back to the MMA warp.
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if constexpr (kCluster == 2) mbar_arrive_cluster(tempty_leader + 8u * as);
        else mbar_arrive(tempty_leader + 8u * as);
      }
    }
    flush_arrivals();
  }

  // ================================ Teardown ================================
  __syncwarp();  // single-lane role loops: reconverge each warp before the aligned barrier
  tc_fence_before();
  if constexpr (kCluster == 2) cluster_sync(); else __syncthreads();
  if (warp_idx == 2) {
    tc_fence_after();
    tmem_dealloc<kCluster>(tmem_base, 512);
  }
  if constexpr (kFuse == 0) {
    if (args.carry.debug != nullptr && blockIdx.x == 0 && threadIdx.x == 0) args.carry.debug[7] = carry_now_ns();
  }
}

// ------------------------------------------------------------------------------------------------------
// Host launcher
// ------------------------------------------------------------------------------------------------------
template <int kCluster, Major kAMajor, Major kBMajor, int kFuse>
static cudaError_t launch_one(const CUtensorMap& ta, const CUtensorMap& tb_, const GemmArgs& args, int num_sms,
                              cudaStream_t stream) {
  using S = GemmSmem<kCluster>;
  auto kern = gemm_bf16_kernel<kCluster, kAMajor, kBMajor, kFuse>;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, S::kTotalCarry);
    if (e != cudaSuccess) return e;
    configured = true;
  }
  const bool carrying = kFuse == 0 && args.carry.slice[0].kind != 0;
  const int num_tiles = args.num_m_tiles * args.num_n_tiles;
  int clusters = num_sms / kCluster;
  if (kFuse == 1) clusters -= args.fuse.comm_clusters;   // copy clusters share their SMs with nobody
  if (clusters > num_tiles) clusters = num_tiles;
  if (clusters < 1) clusters = 1;
  if (kFuse == 1) clusters += args.fuse.comm_clusters;
  if (kFuse == 0 && args.dynamic) clusters = num_tiles;   // one cluster per tile; extra ones are cancelled on device
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(clusters * kCluster);
  cfg.blockDim = dim3(kNumThreads);
  cfg.dynamicSmemBytes = carrying ? S::kTotalCarry : S::kTotal;   // the ring only when there is something to carry
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = kCluster;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kern, ta, tb_, args);
}

// 1 = dynamic (cluster-launch-control) tile scheduling, 0 = static striping.  mode < 0 only queries.
int gemm_sched_mode(int mode) {
  // default static: ~1-3 % faster when the GEMM owns the GPU (profiles/gemm_sched_ab_run21.txt); the sharding engine
  // switches to dynamic when collectives run next to the GEMMs.  TORCHACC_B200_GEMM_SCHED=static|dynamic pins it.
  static int current = [] {
    const char* e = getenv("TORCHACC_B200_GEMM_SCHED");
    return (e && e[0] == 'd') ? 1 : 0;
  }();
  if (mode >= 0) current = mode ? 1 : 0;
  return current;
}

static bool fill_common(GemmArgs& args, void* D, const void* bias, int M, int N, int K, long long ldd, bool out_fp32,
                        const void* C, long long ldc, int cluster) {
  args.D = D;
  args.bias = reinterpret_cast<const __nv_bfloat16*>(bias);
  args.M = M; args.N = N; args.K = K;
  args.ldd = ldd;
  args.C = C; args.ldc = ldc;
  args.out_fp32 = out_fp32 ? 1 : 0;
  const int tile_m = kBlockMCta * cluster;
  args.num_m_tiles = (M + tile_m - 1) / tile_m;
  args.num_n_tiles = (N + kBlockN - 1) / kBlockN;
  args.dynamic = gemm_sched_mode(-1);
  args.fp16 = 0;
  args.swiglu = 0; args.swiglu_F = 0; args.gu_out = nullptr; args.ldgu = 0;
  memset(&args.fuse, 0, sizeof(args.fuse));
  memset(&args.carry, 0, sizeof(args.carry));
  return true;
}

cudaError_t gemm_bf16(const void* A, const void* B, void* D, const void* bias, int M, int N, int K, long long lda,
                      long long ldb, long long ldd, bool a_mn_major, bool b_mn_major, bool out_fp32, bool accumulate,
                      int cluster, int num_sms, cudaStream_t stream, bool is_fp16) {
  return gemm_bf16_ex(A, B, D, bias, accumulate ? D : nullptr, M, N, K, lda, ldb, ldd, ldd, a_mn_major, b_mn_major,
                      out_fp32, cluster, num_sms, stream, is_fp16);
}

cudaError_t gemm_bf16_ex(const void* A, const void* B, void* D, const void* bias, const void* C, int M, int N, int K,
                         long long lda, long long ldb, long long ldd, long long ldc, bool a_mn_major, bool b_mn_major,
                         bool out_fp32, int cluster, int num_sms, cudaStream_t stream, bool is_fp16) {
  if (M <= 0 || N <= 0) return cudaSuccess;
  if (K <= 0) return cudaErrorInvalidValue;
  if (cluster != 1 && cluster != 2) return cudaErrorInvalidValue;
  const int load_n = kBlockN / cluster;
  CUtensorMap ta, tbm;
  try {
    // K-major: matrix [MN][K]; MN-major: matrix [K][MN].  Inner box is always 64 elements = 128 B.
    ta = a_mn_major ? make_map_2d_bf16(A, K, M, lda, 64, kBlockK) : make_map_2d_bf16(A, M, K, lda, kBlockK, kBlockMCta);
    tbm = b_mn_major ? make_map_2d_bf16(B, K, N, ldb, 64, kBlockK) : make_map_2d_bf16(B, N, K, ldb, kBlockK, load_n);
  } catch (const std::exception& e) {
    fprintf(stderr, "%s\n", e.what());
    return cudaErrorInvalidValue;
  }
  GemmArgs args;
  fill_common(args, D, bias, M, N, K, ldd, out_fp32, C, ldc, cluster);
  args.fp16 = is_fp16 ? 1 : 0;
  // pending FSDP collectives ride along: bytes proportional to this GEMM's FLOPs (static scheduling only -- under
  // dynamic scheduling warp 3 is the tile scheduler and collectives run as their own kernels)
  if (!args.dynamic) carry_take(2.0 * (double)M * (double)N * (double)K, &args.carry);

#define TB_DISPATCH(CL, AM, BM) return launch_one<CL, AM, BM, 0>(ta, tbm, args, num_sms, stream)
  if (cluster == 2) {
    if (!a_mn_major && !b_mn_major) TB_DISPATCH(2, Major::K, Major::K);
    if (!a_mn_major && b_mn_major) TB_DISPATCH(2, Major::K, Major::MN);
    if (a_mn_major && !b_mn_major) TB_DISPATCH(2, Major::MN, Major::K);
    TB_DISPATCH(2, Major::MN, Major::MN);
  } else {
    if (!a_mn_major && !b_mn_major) TB_DISPATCH(1, Major::K, Major::K);
    if (!a_mn_major && b_mn_major) TB_DISPATCH(1, Major::K, Major::MN);
    if (a_mn_major && !b_mn_major) TB_DISPATCH(1, Major::MN, Major::K);
    TB_DISPATCH(1, Major::MN, Major::MN);
  }
#undef TB_DISPATCH
}

// h[M][F] = silu(x Wg^T) * (x Wu^T) with W_gu = [Wg; Wu] ([2F][K], K-major); optionally also gu[M][2F] = x W_gu^T.
cudaError_t gemm_swiglu_bf16(const void* A, const void* Wgu, void* H, void* GU, int M, int F, int K, long long lda,
                             long long ldb, long long ldh, long long ldgu, int num_sms, cudaStream_t stream,
                             bool is_fp16) {
  if (M <= 0 || F <= 0) return cudaSuccess;
  if (K <= 0 || F % 128 != 0 || ldh % 8 != 0 || (GU != nullptr && ldgu % 8 != 0)) return cudaErrorInvalidValue;
  CUtensorMap ta, tbm;
  try {
    ta = make_map_2d_bf16(A, M, K, lda, kBlockK, kBlockMCta);
    tbm = make_map_2d_bf16(Wgu, 2 * (uint64_t)F, K, ldb, kBlockK, kBlockN / 2);
  } catch (const std::exception& e) {
    fprintf(stderr, "%s\n", e.what());
    return cudaErrorInvalidValue;
  }
  GemmArgs args;
  fill_common(args, H, nullptr, M, F, K, ldh, false, nullptr, 0, 2);
  args.num_n_tiles = F / 128;           // each 256-wide accumulator tile yields 128 output columns
  args.swiglu = 1; args.swiglu_F = F; args.gu_out = GU; args.ldgu = ldgu;
  args.fp16 = is_fp16 ? 1 : 0;
  args.dynamic = 0;                      // the gate/up row pairing relies on the static pair layout
  carry_take(4.0 * (double)M * (double)F * (double)K, &args.carry);
  return launch_one<2, Major::K, Major::K, 0>(ta, tbm, args, num_sms, stream);
}

// ------------------------------------------------------------------------------------------------------
// Fused tensor-parallel entry points (cta_group::2 kernels only)
// ------------------------------------------------------------------------------------------------------
static cudaError_t check_fused(int world, int rows_per_rank, int N, int channel) {
  if (world < 2 || world > kMaxPeers || channel < 0 || channel >= 64) return cudaErrorInvalidValue;
  if (rows_per_rank % (kBlockMCta * 2) != 0) return cudaErrorInvalidValue;   // M tiles must not straddle ranks
  if (N % 8 != 0) return cudaErrorInvalidValue;
  return cudaSuccess;
}

cudaError_t ag_gemm_bf16(const uint64_t* peer_a_full, const uint64_t* pad_ptrs, void* a_full, const void* B, void* D,
                         const void* bias, int rows_per_rank, int N, int K, long long ldb, long long ldd,
                         bool b_mn_major, int rank, int world, uint32_t* flags, uint32_t flag_target,
                         uint32_t* block_counter, int channel, uint32_t epoch, int comm_clusters, int num_sms,
                         cudaStream_t stream) {
  cudaError_t e = check_fused(world, rows_per_rank, N, channel);
  if (e != cudaSuccess) return e;
  if (K % 8 != 0 || comm_clusters < 1 || comm_clusters * 2 >= num_sms) return cudaErrorInvalidValue;
  const int M = rows_per_rank * world;
  CUtensorMap ta, tbm;
  try {
    ta = make_map_2d_bf16(a_full, M, K, K, kBlockK, kBlockMCta);
    tbm = b_mn_major ? make_map_2d_bf16(B, K, N, ldb, 64, kBlockK) : make_map_2d_bf16(B, N, K, ldb, kBlockK, kBlockN / 2);
  } catch (const std::exception& ex) {
    fprintf(stderr, "%s\n", ex.what());
    return cudaErrorInvalidValue;
  }
  GemmArgs args;
  fill_common(args, D, bias, M, N, K, ldd, false, nullptr, 0, 2);
  FuseArgs& f = args.fuse;
  for (int i = 0; i < world; ++i) {
    f.peer[i] = reinterpret_cast<void*>(peer_a_full[i]);
    f.pads[i] = reinterpret_cast<uint32_t*>(pad_ptrs[i]);
  }
  f.a_full = a_full; f.flags = flags; f.block_counter = block_counter;
  f.rank = rank; f.world = world; f.rows_per_rank = rows_per_rank; f.channel = channel;
  f.epoch = epoch; f.flag_target = flag_target; f.comm_clusters = comm_clusters;
  if (b_mn_major) return launch_one<2, Major::K, Major::MN, 1>(ta, tbm, args, num_sms, stream);
  return launch_one<2, Major::K, Major::K, 1>(ta, tbm, args, num_sms, stream);
}

cudaError_t gemm_rs_bf16(const void* A, const void* B, const uint64_t* peer_stage, const uint64_t* peer_counters,
                         const uint64_t* pad_ptrs, int rows_per_rank, int N, int K, long long lda, long long ldb,
                         bool a_mn_major, bool b_mn_major, int rank, int world, int channel, uint32_t epoch,
                         int num_sms, cudaStream_t stream) {
  cudaError_t e = check_fused(world, rows_per_rank, N, channel);
  if (e != cudaSuccess) return e;
  if (N % 32 != 0) return cudaErrorInvalidValue;
  const int M = rows_per_rank * world;
  CUtensorMap ta, tbm;
  try {
    ta = a_mn_major ? make_map_2d_bf16(A, K, M, lda, 64, kBlockK) : make_map_2d_bf16(A, M, K, lda, kBlockK, kBlockMCta);
    tbm = b_mn_major ? make_map_2d_bf16(B, K, N, ldb, 64, kBlockK) : make_map_2d_bf16(B, N, K, ldb, kBlockK, kBlockN / 2);
  } catch (const std::exception& ex) {
    fprintf(stderr, "%s\n", ex.what());
    return cudaErrorInvalidValue;
  }
  GemmArgs args;
  fill_common(args, nullptr, nullptr, M, N, K, /*ldd=*/N, false, nullptr, 0, 2);
  FuseArgs& f = args.fuse;
  for (int i = 0; i < world; ++i) {
    f.peer[i] = reinterpret_cast<void*>(peer_stage[i]);
    f.peer_cnt[i] = reinterpret_cast<uint32_t*>(peer_counters[i]);
    f.pads[i] = reinterpret_cast<uint32_t*>(pad_ptrs[i]);
  }
  f.rank = rank; f.world = world; f.rows_per_rank = rows_per_rank; f.channel = channel; f.epoch = epoch;
  f.slot_stride = (long long)rows_per_rank * ((N + kBlockN - 1) / kBlockN * kBlockN);   // block-major, padded N
#define TB_RS(AM, BM) return launch_one<2, AM, BM, 2>(ta, tbm, args, num_sms, stream)
  if (!a_mn_major && !b_mn_major) TB_RS(Major::K, Major::K);
  if (!a_mn_major && b_mn_major) TB_RS(Major::K, Major::MN);
  if (a_mn_major && !b_mn_major) TB_RS(Major::MN, Major::K);
  TB_RS(Major::MN, Major::MN);
#undef TB_RS
}

// out[rows, N] = sum_src stage[src] (+ residual), once every source has delivered `expected` arrivals.
// Staging slots are block-major (see the kFuse == 2 epilogue): one warp handles one 2 KB block = 32 rows x 32 columns;
// its four loads per source are 512-byte coalesced, each lane then owns 64 contiguous output bytes of one row.
__global__ void __launch_bounds__(256)
rs_reduce_kernel(const __nv_bfloat16* __restrict__ stage, const uint32_t* __restrict__ counters, uint32_t expected,
                 const __nv_bfloat16* __restrict__ residual, __nv_bfloat16* __restrict__ out, int rows, int N, int world,
                 long long slot_stride) {
  if (threadIdx.x < world)
    spin_until_count(counters + threadIdx.x, expected, -1, (int)threadIdx.x, "GEMM -> reduce-scatter arrivals");
  __syncthreads();
  const int num_n_tiles = (N + 255) / 256;
  const long long num_blocks = (long long)(rows / 256) * num_n_tiles * 64;
  const int lane = threadIdx.x & 31;
  const long long warps = (long long)gridDim.x * (blockDim.x >> 5);
  for (long long w = blockIdx.x * (long long)(blockDim.x >> 5) + (threadIdx.x >> 5); w < num_blocks; w += warps) {
    const int c = (int)(w & 7), q = (int)((w >> 3) & 3), cta = (int)((w >> 5) & 1);
    const long long tile = w >> 6;
    const int tn = (int)(tile % num_n_tiles);
    const long long tm = tile / num_n_tiles;
    const int col = tn * 256 + c * 32;
    if (col + 32 > N) continue;
    const long long row = tm * 256 + cta * 128 + q * 32 + lane;
    float acc[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) acc[j] = 0.f;
    for (int s = 0; s < world; ++s) {
      const uint4* src = reinterpret_cast<const uint4*>(stage + s * slot_stride) + w * 128 + lane;
      uint4 u[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) u[j] = ld_volatile_v4(src + j * 32);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float2 a = unpack_bf16x2(u[j].x), b = unpack_bf16x2(u[j].y), cc = unpack_bf16x2(u[j].z), d = unpack_bf16x2(u[j].w);
        acc[8 * j + 0] += a.x; acc[8 * j + 1] += a.y; acc[8 * j + 2] += b.x; acc[8 * j + 3] += b.y;
        acc[8 * j + 4] += cc.x; acc[8 * j + 5] += cc.y; acc[8 * j + 6] += d.x; acc[8 * j + 7] += d.y;
      }
    }
    const long long o_off = row * N + col;
    if (residual) {
      const uint4* rp = reinterpret_cast<const uint4*>(residual + o_off);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const uint4 u = rp[j];
        float2 a = unpack_bf16x2(u.x), b = unpack_bf16x2(u.y), cc = unpack_bf16x2(u.z), d = unpack_bf16x2(u.w);
        acc[8 * j + 0] += a.x; acc[8 * j + 1] += a.y; acc[8 * j + 2] += b.x; acc[8 * j + 3] += b.y;
        acc[8 * j + 4] += cc.x; acc[8 * j + 5] += cc.y; acc[8 * j + 6] += d.x; acc[8 * j + 7] += d.y;
      }
    }
    uint4* op = reinterpret_cast<uint4*>(out + o_off);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      uint4 o;
      o.x = pack_bf16x2(acc[8 * j + 0], acc[8 * j + 1]); o.y = pack_bf16x2(acc[8 * j + 2], acc[8 * j + 3]);
      o.z = pack_bf16x2(acc[8 * j + 4], acc[8 * j + 5]); o.w = pack_bf16x2(acc[8 * j + 6], acc[8 * j + 7]);
      op[j] = o;
    }
  }
}

cudaError_t rs_reduce_bf16(const void* stage, const uint32_t* counters, uint32_t expected, const void* residual,
                           void* out, int rows, int N, int world, long long slot_stride, int num_sms,
                           cudaStream_t stream) {
  if (N % 32 != 0 || rows % 256 != 0) return cudaErrorInvalidValue;
  long long blocks = (long long)(rows / 256) * ((N + 255) / 256) * 64 / 8;   // 8 warps per CTA
  if (blocks > (long long)num_sms * 4) blocks = (long long)num_sms * 4;
  if (blocks < 1) blocks = 1;
  rs_reduce_kernel<<<(unsigned)blocks, 256, 0, stream>>>((const __nv_bfloat16*)stage, counters, expected,
                                                          (const __nv_bfloat16*)residual, (__nv_bfloat16*)out, rows, N,
                                                          world, slot_stride);
  return cudaGetLastError();
}

}  // namespace tb
