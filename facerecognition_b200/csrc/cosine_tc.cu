// cosine_tc.cu — K1 (bf16): similarity GEMM on tcgen05 tensor cores with a fused per-query top-k.
//
// Replaces, for bf16 galleries, the np.dot(E, P.T) + argsort of the reference's batched evaluation
// (notebooks/evaluate_arcface_kaggle.ipynb:618,713), faiss.IndexFlatIP.search
// (inference/recognition_engine.py:304) and the per-identity loops (recognition_engine.py:277-289,
// web_app.py:545-554).  The Q x N score matrix lives only in TMEM: it never reaches shared or global
// memory.
//
// Shape.  D[128 queries x 256 gallery rows] += A[128 x 16] * B[256 x 16]^T, bf16 in, fp32 accumulate,
// one tcgen05.mma per 16 of K, issued by one thread.  Queries sit on the M axis so that TMEM lane ==
// query: an epilogue thread owns one query and streams that query's scores out of TMEM with
// tcgen05.ld, keeping its best-k list privately (no cross-thread traffic for the selection).
//
// Work unit = (128-query tile qt, group of G consecutive 256-row gallery tiles); unit id u = group * n_qt
// + qt, CTA p takes u = p, p + grid, ...  Units running at the same time therefore share their gallery
// group, which is read from HBM once and from L2 by the other query tiles.
//
// Warp roles (256 threads):  warp 0 = TMA producer (one lane), warp 1 = MMA issuer (one lane),
// warp 2 = TMEM allocator, warps 4-7 = epilogue (TMEM lanes 32*(w%4)..+31).
// Shared memory: A = the unit's 128 x dim query tile (dim/64 k-blocks of 16 KB, resident for the unit),
// B = ring of 32 KB stages (256 rows x 64 of K), both 128-byte swizzled K-major as written by TMA.
// TMEM: 2 accumulator stages x 256 columns, so the epilogue of tile t overlaps the MMAs of tile t+1.
//
// Two kernels share the epilogue code: cosine_tc_kernel (one CTA per unit, described above; batches of up to 128
// queries) and cosine_tc_pair_kernel (tcgen05.mma.cta_group::2: two CTAs of a cluster share every MMA and each loads
// half of the gallery tile; every larger batch, all list lengths).  Operands are bf16 or fp16 (launch parameter).
#include "frb_common.cuh"

#include <stdlib.h>

#include "tc_ptx.cuh"

namespace frb {

constexpr int kTcBlockM = 128;   // queries per tile (TMEM lanes)
constexpr int kTcBlockN = 256;   // gallery rows per tile (TMEM columns per accumulator stage)
constexpr int kTcBlockK = 64;    // bf16 elements per 128-byte swizzle row
constexpr int kTcUmmaK = 16;
constexpr int kTcThreads = 256;
constexpr int kTcMaxKBlocks = 8;  // dim <= 512
constexpr uint32_t kTcABytesPerKb = kTcBlockM * kTcBlockK * 2;  // 16 KB
constexpr uint32_t kTcBBytesPerStage = kTcBlockN * kTcBlockK * 2;  // 32 KB
constexpr int kTcMaxStages = 8;
constexpr int kTcPrefetchDist = 6;  // gallery tiles (256 rows = 256 KB) kept ahead of the TMA loads in L2, at most
// L2 the prefetched tiles of all concurrently walked gallery groups may occupy.  With few query tiles per group the
// grid walks many groups at once (148 of them for <= 128 queries): 6 tiles ahead in each is 227 MB, which evicts lines
// before their demand load arrives -- measured 1.78x the gallery in DRAM reads and 3.9 TB/s instead of 6.1 TB/s
// (profiles/r1_tc_prefetch_sweep.txt).  The distance shrinks to fit this budget; 0 turns the prefetch off.
constexpr size_t kTcPrefetchL2Budget = 8u << 20;
constexpr int kTcShareMinK = 8;          // lists longer than this exchange share-of-k bounds (TcParams::share)
constexpr int kTcShareMaxGroups = 64;    // ... when a pass has at most this many gallery groups (one coalesced load each per tile)
constexpr size_t kTcSmemLimit = 227 * 1024;

struct TcBarriers {
    uint64_t full[kTcMaxStages];
    uint64_t empty[kTcMaxStages];
    uint64_t a_full, a_empty;
    uint64_t tmem_full[2], tmem_empty[2];
    uint32_t tmem_base;
};

// kind::f16 instruction descriptor: D=f32 (1<<4), A and B bf16 (1<<7, 1<<10) or fp16 (format 0), both K-major, N>>3 at 17, M>>4 at 24
constexpr uint32_t kTcIdescF16 = (1u << 4) | ((uint32_t)(kTcBlockN >> 3) << 17) | ((uint32_t)(kTcBlockM >> 4) << 24);
constexpr uint32_t kTcIdescBf16 = kTcIdescF16 | (1u << 7) | (1u << 10);


// ---- register-resident best-k list, RK = 8 / 16 / 32 / 64 slots (k <= RK; 8 covers the common top-1 / top-5) ---
// s[] stays sorted descending over all RK slots; an element is admitted when it beats s[k-1]
// (strictly, so an equal score with a later row stays behind), dropped into the last slot and bubbled
// up with predicated swaps: no local memory, ~6 instructions per slot, executed only on admissions.
constexpr int kTcMaxRegK = 64;  // == FRB_MAX_K: every list fits the register file (the local-memory list, RK = 0, is kept as the generic form)

template <int RK>
__device__ __forceinline__ float reg_kth(const float (&s)[RK], int k)
{
    float r = s[0];
#pragma unroll
    for (int j = 1; j < RK; j++) r = (k == j + 1) ? s[j] : r;
    return r;
}

template <int RK>
__device__ __forceinline__ void reg_insert(float (&s)[RK], int (&id)[RK], float v, int idx)
{
    s[RK - 1] = v;
    id[RK - 1] = idx;
#pragma unroll
    for (int j = RK - 1; j > 0; --j) {
        const bool sw = s[j] > s[j - 1];
        const float ts = s[j - 1];
        const int ti = id[j - 1];
        s[j - 1] = sw ? s[j] : ts;
        id[j - 1] = sw ? id[j] : ti;
        s[j] = sw ? ts : s[j];
        id[j] = sw ? ti : id[j];
    }
}

// Long lists: admitted (score, column) pairs wait in a per-lane queue and are inserted by the whole warp at once
// (see the epilogue).  A lane without an r-th pending entry re-inserts its own last slot, which is a no-op.
constexpr int kTcQueue = 4;
template <int RS>
__device__ __forceinline__ void drain_queue(float (&rs)[RS], int (&ri)[RS], const float (&qv)[kTcQueue], const int (&qi)[kTcQueue],
                                            int &qn)
{
#pragma unroll
    for (int r = 0; r < kTcQueue; r++) {
        if (!__any_sync(0xffffffffu, r < qn)) break;
        // A queued score was admitted against the threshold of the LAST drain; earlier inserts of this drain may have
        // raised the list's last slot above it.  reg_insert overwrites slot RS-1 unconditionally, so such an entry must
        // be dropped here (strictly: an equal score from a later row stays behind) or, with k == RS, it would evict a
        // row that belongs in the top-k.
        const bool has = r < qn && qv[r] > rs[RS - 1];
        reg_insert<RS>(rs, ri, has ? qv[r] : rs[RS - 1], has ? qi[r] : ri[RS - 1]);
    }
    qn = 0;
}

// v[j] for a run-time j without spilling v[] to local memory: 5-level select tree (31 SEL)
__device__ __forceinline__ float select32(const float *v, int j)
{
    float a[16];
#pragma unroll
    for (int i = 0; i < 16; i++) a[i] = (j & 16) ? v[i + 16] : v[i];
#pragma unroll
    for (int w = 8; w > 0; w >>= 1)
#pragma unroll
        for (int i = 0; i < w; i++) a[i] = (j & w) ? a[i + w] : a[i];
    return a[0];
}

__device__ __forceinline__ float max32(const float *v)
{
    float m[16];
#pragma unroll
    for (int j = 0; j < 16; j++) m[j] = fmaxf(v[j], v[j + 16]);
#pragma unroll
    for (int w = 8; w > 0; w >>= 1)
#pragma unroll
        for (int j = 0; j < w; j++) m[j] = fmaxf(m[j], m[j + w]);
    return m[0];
}

// ---- cross-CTA admission threshold ---------------------------------------------------------------
// thr[q] holds a lower bound on query q's final k-th best score, stored one ulp BELOW a score that k
// rows are already known to reach (so an equal score from another gallery group still gets in and the
// lowest-row tie order is preserved).  Every unit starts from it instead of from -inf; without it each
// unit re-learns its threshold and, with 32 independent lists per warp, nearly every 32-column chunk
// takes the (slow, divergent) admission path.  It only ever discards rows that cannot be in the
// final top-k, so results do not depend on timing.
__device__ __forceinline__ float next_below(float x)
{
    if (x == -INFINITY) return x;
    if (x == 0.f) return -1.17549435e-38f;
    const int b = __float_as_int(x);
    return __int_as_float(x > 0.f ? b - 1 : b + 1);
}
__device__ __forceinline__ void atomic_max_f32(float *addr, float v)
{
    if (v >= 0.f)
        atomicMax(reinterpret_cast<int *>(addr), __float_as_int(v));
    else
        atomicMin(reinterpret_cast<unsigned int *>(addr), __float_as_uint(v));
}
__device__ __forceinline__ void st_relaxed_f32(float *addr, float v)
{
    asm volatile("st.relaxed.gpu.global.f32 [%0], %1;" ::"l"(addr), "f"(v) : "memory");
}
__device__ __forceinline__ float ld_relaxed_f32(const float *addr)
{
    float v;
    asm volatile("ld.relaxed.gpu.global.f32 %0, [%1];" : "=f"(v) : "l"(addr));
    return v;
}

struct TcParams {
    int64_t n_query, n_gallery;
    int k_blocks;          // dim / 64
    uint32_t idesc;        // UMMA instruction descriptor: bf16 or fp16 operands
    int stages;            // B ring depth
    int64_t n_qtiles, tiles_per_group, n_groups;  // units of THIS launch = n_qtiles * n_groups
    int64_t tile_begin, tile_end;                 // gallery tiles [tile_begin, tile_end) are scanned by this launch
    int k;
    int prefetch_dist;     // gallery tiles prefetched into L2 ahead of the TMA loads (0 = off)
    int64_t idx_base;
    float *thr;            // [n_query] shared admission thresholds, -inf on entry
    // Long lists (k > 8) only.  share[g * n_query + q] = one ulp below the share_rank-th best score that the unit
    // (q's tile, gallery group g) has seen, share_rank = ceil(k / n_groups); bytes 0xFF = not published yet.  The
    // groups are disjoint row ranges, so the minimum over g is reached by n_groups * share_rank >= k rows: a valid
    // lower bound on the final k-th best that is as tight as ONE list over all rows seen so far, whereas the maximum
    // of the units' own k-th bests (thr) is only as tight as a list over 1/n_groups of them.  NULL: not used.
    float *share;
    int share_rank;
    int *cand_cnt;         // [n_query] candidates appended so far, 0 on entry
    int64_t cand_cap;      // slots per query = (total groups) * k: every unit can always append its whole list
    float *cand_scores;    // [n_query, cand_cap] unordered
    int64_t *cand_idx;
};

// ---- epilogue pieces shared by the single-CTA and the CTA-pair kernel (register-resident lists) -------------------
// 32 scores of one query (TMEM columns id0 .. id0 + 31 of the unit) against its list.  QUEUED (lists of 16+ slots): a
// sorted insert is ~6 RS instructions and, done where the candidate is found, the whole warp pays for ONE lane's insert.
// Lanes park their admitted scores in a 4-slot queue instead; when any lane's queue is full every lane drains its own,
// so one pass of the insert code serves up to 32 candidates.  Thresholds move only at a drain: a score admitted against
// a stale threshold is re-checked against the list's last slot inside drain_queue and dropped there.
template <int RS, bool QUEUED>
__device__ __forceinline__ void tc_admit32(const float *v, int id0, int k, float gthr, float &adm, float &kth, float (&rs)[RS],
                                           int (&ri)[RS], float (&qv)[kTcQueue], int (&qi)[kTcQueue], int &qn)
{
    if (QUEUED) {
        if (__any_sync(0xffffffffu, max32(v) > adm)) {
            uint32_t cand = 0;
#pragma unroll
            for (int j = 0; j < 32; j++) cand |= (v[j] > adm) ? (1u << j) : 0u;
            while (__any_sync(0xffffffffu, cand != 0)) {
                if (cand) {
                    const int j = __ffs(cand) - 1;
                    cand &= cand - 1;
                    const float x = select32(v, j);
                    if (x > adm) {
                        const int id = id0 + j;
#pragma unroll
                        for (int r = 0; r < kTcQueue; r++) {
                            qv[r] = (qn == r) ? x : qv[r];
                            qi[r] = (qn == r) ? id : qi[r];
                        }
                        qn++;
                    }
                }
                if (__any_sync(0xffffffffu, qn == kTcQueue)) {
                    drain_queue<RS>(rs, ri, qv, qi, qn);
                    kth = reg_kth<RS>(rs, k);
                    adm = fmaxf(kth, gthr);
                }
            }
        }
    } else if (max32(v) > adm) {   // rare once the thresholds have warmed up
        // candidate mask first (straight-line), then visit this lane's candidates in column order; the warp
        // iterates max-popcount times instead of walking 32 branchy checks
        uint32_t cand = 0;
#pragma unroll
        for (int j = 0; j < 32; j++) cand |= (v[j] > adm) ? (1u << j) : 0u;
        while (cand) {
            const int j = __ffs(cand) - 1;
            cand &= cand - 1;
            const float x = select32(v, j);
            if (x > adm) {
                reg_insert<RS>(rs, ri, x, id0 + j);
                kth = reg_kth<RS>(rs, k);
                adm = fmaxf(kth, gthr);
            }
        }
    }
}

// Once per tile: publish this unit's bound for query q if it improved (thr[q], and for long lists its share-of-k
// bound), pick up the other units'.  kth: the unit's own k-th best.
template <int RS>
__device__ __forceinline__ void tc_exchange_bounds(const TcParams &p, int64_t q, int64_t grp, const float (&rs)[RS], float kth,
                                                   float &published, float &pub_share, float &gthr, float &adm)
{
    float mine = next_below(kth);
    if (RS >= 16 && p.share) {
        float *slot = p.share + q;   // [group][query]: a warp's 32 queries share sectors
        const float part = next_below(reg_kth<RS>(rs, p.share_rank));
        if (part > pub_share) {
            st_relaxed_f32(slot + grp * p.n_query, part);
            pub_share = part;
        }
        float m = INFINITY;
        for (int64_t g = 0; g < p.n_groups; g++) {
            const float v = ld_relaxed_f32(slot + g * p.n_query);
            m = (__float_as_uint(v) == 0xFFFFFFFFu) ? -INFINITY : fminf(m, v);
        }
        mine = fmaxf(mine, m);
    }
    if (mine > published) {
        atomic_max_f32(p.thr + q, mine);
        published = mine;
    }
    gthr = fmaxf(gthr, ld_relaxed_f32(p.thr + q));
    adm = fmaxf(kth, gthr);
}

// RK: slots of the register-resident list (8, 16, 32, 64), or 0 for the local-memory list
template <int RK>
__global__ void __launch_bounds__(kTcThreads, 1)
cosine_tc_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_g,
                 const __grid_constant__ CUtensorMap tmap_pf, const TcParams p)
{
    extern __shared__ unsigned char smem_raw[];
    // SWIZZLE_128B atoms need 1024-byte alignment
    unsigned char *smem = reinterpret_cast<unsigned char *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    unsigned char *smem_a = smem;
    unsigned char *smem_b = smem + (size_t)p.k_blocks * kTcABytesPerKb;
    TcBarriers *bars = reinterpret_cast<TcBarriers *>(smem_b + (size_t)p.stages * kTcBBytesPerStage);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t n_units = p.n_qtiles * p.n_groups;

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_q) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_g) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_pf) : "memory");
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < p.stages; s++) {
            mbar_init(&bars->full[s], 1);
            mbar_init(&bars->empty[s], 1);
        }
        mbar_init(&bars->a_full, 1);
        mbar_init(&bars->a_empty, 1);
        for (int s = 0; s < 2; s++) {
            mbar_init(&bars->tmem_full[s], 1);
            mbar_init(&bars->tmem_empty[s], 4);  // one arrive per epilogue warp
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&bars->tmem_base)), "r"(512u)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = bars->tmem_base;

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0, a_phase = 0;
            for (int64_t u = blockIdx.x; u < n_units; u += gridDim.x) {
                const int64_t qt = u % p.n_qtiles, grp = u / p.n_qtiles;
                const int64_t t0 = p.tile_begin + grp * p.tiles_per_group;
                const int64_t t1 = (t0 + p.tiles_per_group < p.tile_end) ? t0 + p.tiles_per_group : p.tile_end;
                mbar_wait(&bars->a_empty, a_phase ^ 1);  // previous unit's MMAs have finished reading A
                mbar_expect_tx(&bars->a_full, (uint32_t)p.k_blocks * kTcABytesPerKb);
                for (int kb = 0; kb < p.k_blocks; kb++)
                    tma_load_2d(smem_a + (size_t)kb * kTcABytesPerKb, &tmap_q, &bars->a_full, kb * kTcBlockK, (int)(qt * kTcBlockM));
                a_phase ^= 1;
                // L2 prefetch.  The query tiles that share this gallery group walk it in lockstep, so every
                // line's first touch would pay DRAM latency in all of them at once; instead tile t + p.prefetch_dist
                // is pulled into L2 ahead of time, the duty split round-robin over the sharing query tiles.
                for (int64_t t = t0 + qt; t < t0 + p.prefetch_dist && t < t1; t += p.n_qtiles) {
                    for (int c = 0; c < p.k_blocks * kTcBlockK; c += 256) tma_prefetch_l2_2d(&tmap_pf, c, (int)(t * kTcBlockN));
                }
                for (int64_t t = t0; t < t1; t++) {
                    const int64_t tp = t + p.prefetch_dist;
                    if (tp < t1 && (tp - t0) % p.n_qtiles == qt) {
                        for (int c = 0; c < p.k_blocks * kTcBlockK; c += 256) tma_prefetch_l2_2d(&tmap_pf, c, (int)(tp * kTcBlockN));
                    }
                    for (int kb = 0; kb < p.k_blocks; kb++) {
                        mbar_wait(&bars->empty[stage], phase ^ 1);
                        mbar_expect_tx(&bars->full[stage], kTcBBytesPerStage);
                        tma_load_2d(smem_b + (size_t)stage * kTcBBytesPerStage, &tmap_g, &bars->full[stage], kb * kTcBlockK,
                                    (int)(t * kTcBlockN));
                        if (++stage == p.stages) { stage = 0; phase ^= 1; }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        if (lane == 0) {
            int stage = 0, acc = 0;
            uint32_t phase = 0, a_phase = 0, acc_phase = 0;
            for (int64_t u = blockIdx.x; u < n_units; u += gridDim.x) {
                const int64_t grp = u / p.n_qtiles;
                const int64_t t0 = p.tile_begin + grp * p.tiles_per_group;
                const int64_t t1 = (t0 + p.tiles_per_group < p.tile_end) ? t0 + p.tiles_per_group : p.tile_end;
                mbar_wait(&bars->a_full, a_phase);
                a_phase ^= 1;
                tcgen05_fence_after();
                for (int64_t t = t0; t < t1; t++) {
                    mbar_wait(&bars->tmem_empty[acc], acc_phase ^ 1);  // epilogue drained this accumulator
                    tcgen05_fence_after();
                    const uint32_t tmem_d = tmem_base + (uint32_t)acc * kTcBlockN;
                    for (int kb = 0; kb < p.k_blocks; kb++) {
                        mbar_wait(&bars->full[stage], phase);
                        tcgen05_fence_after();
                        const uint64_t da = make_sw128_desc(smem_u32(smem_a + (size_t)kb * kTcABytesPerKb));
                        const uint64_t db = make_sw128_desc(smem_u32(smem_b + (size_t)stage * kTcBBytesPerStage));
#pragma unroll
                        for (int k4 = 0; k4 < kTcBlockK / kTcUmmaK; k4++) {
                            // advance 16 elements (32 B) along K inside the 128-byte swizzle row: +2 in 16-byte units
                            umma_bf16(tmem_d, da + (uint64_t)(k4 * 2), db + (uint64_t)(k4 * 2), p.idesc, (kb | k4) != 0 ? 1u : 0u);
                        }
                        tcgen05_commit(&bars->empty[stage]);  // stage reusable once these MMAs retire
                        if (++stage == p.stages) { stage = 0; phase ^= 1; }
                    }
                    tcgen05_commit(&bars->tmem_full[acc]);  // accumulator ready for the epilogue
                    if (++acc == 2) { acc = 0; acc_phase ^= 1; }
                }
                tcgen05_commit(&bars->a_empty);  // A may be overwritten by the next unit
            }
        }
    } else if (warp >= 4) {
        // ===================== epilogue: TMEM -> registers -> per-query best-k =====================
        const int ew = warp & 3;                       // TMEM lane quadrant this warp may read
        const int row = ew * 32 + lane;                // query row inside the tile == TMEM lane
        int acc = 0;
        uint32_t acc_phase = 0;
        // REG_LIST: k <= RK, list in registers with 32-bit row offsets relative to the unit's first row;
        // otherwise a local-memory list (k up to FRB_MAX_K), touched only on admissions.
        constexpr bool REG_LIST = RK > 0;
        constexpr int RS = REG_LIST ? RK : 1;
        constexpr bool QUEUED = RK >= 16;   // long register lists take admissions through a per-lane queue
        float qv[kTcQueue];
        int qi[kTcQueue];
        int qn = 0;
#pragma unroll
        for (int r = 0; r < kTcQueue; r++) { qv[r] = -INFINITY; qi[r] = -1; }
        float rs[RS];
        int ri[RS];
        float best_s[REG_LIST ? 1 : FRB_MAX_K];
        int64_t best_i[REG_LIST ? 1 : FRB_MAX_K];
        for (int64_t u = blockIdx.x; u < n_units; u += gridDim.x) {
            const int64_t qt = u % p.n_qtiles, grp = u / p.n_qtiles;
            const int64_t t0 = p.tile_begin + grp * p.tiles_per_group;
            const int64_t t1 = (t0 + p.tiles_per_group < p.tile_end) ? t0 + p.tiles_per_group : p.tile_end;
            const int64_t unit_n0 = t0 * kTcBlockN;
            if (REG_LIST) {
#pragma unroll
                for (int j = 0; j < RS; j++) { rs[j] = -INFINITY; ri[j] = -1; }
            } else {
                list_init<true>(best_s, best_i, p.k);
            }
            const int64_t q = qt * kTcBlockM + row;
            const bool q_live = q < p.n_query;
            float kth = -INFINITY;                                   // this unit's own k-th best
            float gthr = q_live ? ld_relaxed_f32(p.thr + q) : INFINITY;  // dead rows admit nothing
            float adm = gthr;                                        // admit v > adm = max(kth, gthr)
            float published = gthr;
            float pub_share = -INFINITY;
            for (int64_t t = t0; t < t1; t++) {
                mbar_wait(&bars->tmem_full[acc], acc_phase);
                tcgen05_fence_after();
                const int64_t n0 = t * kTcBlockN;
                const int valid = (int)((p.n_gallery - n0) < kTcBlockN ? (p.n_gallery - n0) : kTcBlockN);
                const int col0 = (int)(n0 - unit_n0);
                const uint32_t taddr = tmem_base + ((uint32_t)(ew * 32) << 16) + (uint32_t)acc * kTcBlockN;
#pragma unroll 1
                for (int c0 = 0; c0 < kTcBlockN; c0 += 32) {
                    float v[32];
                    tmem_ld_32x32(taddr + (uint32_t)c0, v);
                    if (c0 + 32 > valid) {  // ragged last tile: TMA zero-filled rows must not compete
#pragma unroll
                        for (int j = 0; j < 32; j++) v[j] = (c0 + j < valid) ? v[j] : -INFINITY;
                    }
                    if (REG_LIST) {
                        tc_admit32<RS, QUEUED>(v, col0 + c0, p.k, gthr, adm, kth, rs, ri, qv, qi, qn);
                    } else if (max32(v) > adm) {   // local-memory list (k beyond the register variants)
                        uint32_t cand = 0;
#pragma unroll
                        for (int j = 0; j < 32; j++) cand |= (v[j] > adm) ? (1u << j) : 0u;
                        while (cand) {
                            const int j = __ffs(cand) - 1;
                            cand &= cand - 1;
                            const float x = select32(v, j);
                            if (x > adm) {
                                kth = list_insert_stream<true>(best_s, best_i, p.k, x, p.idx_base + n0 + c0 + j);
                                adm = fmaxf(kth, gthr);
                            }
                        }
                    }
                }
                tcgen05_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&bars->tmem_empty[acc]);
                if (++acc == 2) { acc = 0; acc_phase ^= 1; }
                // exchange thresholds once per tile: publish ours if it improved, pick up the others'
                if (q_live) tc_exchange_bounds<RS>(p, q, grp, rs, kth, published, pub_share, gthr, adm);
            }
            if (QUEUED && __any_sync(0xffffffffu, qn > 0)) drain_queue<RS>(rs, ri, qv, qi, qn);
            if (q_live) {
                // append this unit's admitted rows to the query's compact candidate buffer
                int nv = 0;
                if (REG_LIST) {
#pragma unroll
                    for (int j = 0; j < RS; j++) nv += (j < p.k && ri[j] >= 0) ? 1 : 0;
                } else {
                    for (int j = 0; j < p.k; j++) nv += best_i[j] >= 0 ? 1 : 0;
                }
                if (nv > 0) {
                    const int64_t o = q * p.cand_cap + atomicAdd(p.cand_cnt + q, nv);
                    if (REG_LIST) {
#pragma unroll
                        for (int j = 0; j < RS; j++)
                            if (j < nv) {
                                p.cand_scores[o + j] = rs[j];
                                p.cand_idx[o + j] = p.idx_base + unit_n0 + ri[j];
                            }
                    } else {
                        for (int j = 0; j < nv; j++) {
                            p.cand_scores[o + j] = best_s[j];
                            p.cand_idx[o + j] = best_i[j];
                        }
                    }
                }
            }
        }
    }

    tcgen05_fence_before();
    __syncthreads();
    if (warp == 2) {
        tcgen05_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
    }
}

// ---- CTA-pair variant (cta_group::2), k <= 8 ---------------------------------------------------------------------
// Two CTAs of a cluster (same TPC) run ONE tcgen05.mma of M = 256: CTA r of the pair owns query tile 2 * qp + r (its
// 128 TMEM lanes, its own epilogue) and loads only rows [r * 128, r * 128 + 128) of every 256-row gallery tile; the
// tensor core reads the other half from the peer's shared memory.  Per CTA that halves the B bytes fetched from L2 and
// held in shared memory: six 16 KB stages fit where the single-CTA kernel has three 32 KB ones, so twice as many
// k-blocks are in flight under the same TMA latency.  The leader (cluster rank 0) issues the MMAs; its full barriers
// count the bytes of BOTH CTAs' TMA loads (cp.async.bulk.tensor ... cta_group::2), its commits are multicast to the
// empty / tmem_full barriers of both CTAs, and the peer's epilogue warps arrive on the leader's tmem_empty barriers.
// Work unit = (pair of query tiles qp, gallery group); pair c of the grid takes units c, c + n_pairs_in_grid, ...
constexpr int kTcPairStagesMax = 6;
constexpr uint32_t kTcPairBBytesPerStage = (kTcBlockN / 2) * kTcBlockK * 2;   // 16 KB: this CTA's half of the gallery tile
constexpr uint32_t kTcIdescPairF16 = (1u << 4) | ((uint32_t)(kTcBlockN >> 3) << 17) | ((uint32_t)((2 * kTcBlockM) >> 4) << 24);
constexpr uint32_t kTcIdescPairBf16 = kTcIdescPairF16 | (1u << 7) | (1u << 10);

struct TcPairBarriers {
    uint64_t full[kTcPairStagesMax];    // leader's copy counts: one arrive.expect_tx (leader) + the bytes of both CTAs
    uint64_t empty[kTcPairStagesMax];   // each CTA's own copy: multicast commit of the MMAs that read the stage
    uint64_t a_full, a_empty;
    uint64_t tmem_full[2], tmem_empty[2];
    uint32_t tmem_base;
};

template <int RS>   // slots of the register-resident best-k list: 8, 16 (the exact-fp32 path's first pass), 32 or 64
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kTcThreads, 1)
cosine_tc_pair_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_gh,
                      const __grid_constant__ CUtensorMap tmap_pf, const TcParams p)
{
    extern __shared__ unsigned char smem_raw[];
    unsigned char *smem = reinterpret_cast<unsigned char *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    unsigned char *smem_a = smem;
    unsigned char *smem_b = smem + (size_t)p.k_blocks * kTcABytesPerKb;
    TcPairBarriers *bars = reinterpret_cast<TcPairBarriers *>(smem_b + (size_t)p.stages * kTcPairBBytesPerStage);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const bool leader = rank == 0;
    const int64_t n_qpairs = (p.n_qtiles + 1) / 2;
    const int64_t n_units = n_qpairs * p.n_groups;
    const int64_t pair0 = blockIdx.x >> 1, pair_step = gridDim.x >> 1;

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_q) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_gh) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_pf) : "memory");
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < p.stages; s++) {
            mbar_init(&bars->full[s], 1);
            mbar_init(&bars->empty[s], 1);
        }
        mbar_init(&bars->a_full, 1);
        mbar_init(&bars->a_empty, 1);
        for (int s = 0; s < 2; s++) {
            mbar_init(&bars->tmem_full[s], 1);
            mbar_init(&bars->tmem_empty[s], 8);   // four epilogue warps of each CTA of the pair
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&bars->tmem_base)), "r"(512u)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    tcgen05_fence_before();
    cluster_sync_all();                       // both CTAs' barriers exist before anyone signals across the pair
    tcgen05_fence_after();
    const uint32_t tmem_base = bars->tmem_base;

    if (warp == 0) {
        // ===================== TMA producer (both CTAs) =====================
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0, a_phase = 0;
            for (int64_t u = pair0; u < n_units; u += pair_step) {
                const int64_t qp = u % n_qpairs, grp = u / n_qpairs;
                const int64_t qt = 2 * qp + rank;     // may be one past the last tile: TMA zero-fills, the epilogue ignores it
                const int64_t t0 = p.tile_begin + grp * p.tiles_per_group;
                const int64_t t1 = (t0 + p.tiles_per_group < p.tile_end) ? t0 + p.tiles_per_group : p.tile_end;
                mbar_wait(&bars->a_empty, a_phase ^ 1);
                if (leader) mbar_expect_tx(&bars->a_full, 2u * (uint32_t)p.k_blocks * kTcABytesPerKb);
                for (int kb = 0; kb < p.k_blocks; kb++)
                    tma_load_2d_pair(smem_a + (size_t)kb * kTcABytesPerKb, &tmap_q, &bars->a_full, kb * kTcBlockK, (int)(qt * kTcBlockM));
                a_phase ^= 1;
                if (leader) {
                    for (int64_t t = t0 + qp; t < t0 + p.prefetch_dist && t < t1; t += n_qpairs)
                        for (int c = 0; c < p.k_blocks * kTcBlockK; c += 256) tma_prefetch_l2_2d(&tmap_pf, c, (int)(t * kTcBlockN));
                }
                for (int64_t t = t0; t < t1; t++) {
                    const int64_t tp = t + p.prefetch_dist;
                    if (leader && tp < t1 && (tp - t0) % n_qpairs == qp)
                        for (int c = 0; c < p.k_blocks * kTcBlockK; c += 256) tma_prefetch_l2_2d(&tmap_pf, c, (int)(tp * kTcBlockN));
                    for (int kb = 0; kb < p.k_blocks; kb++) {
                        mbar_wait(&bars->empty[stage], phase ^ 1);
                        if (leader) mbar_expect_tx(&bars->full[stage], 2u * kTcPairBBytesPerStage);
                        tma_load_2d_pair(smem_b + (size_t)stage * kTcPairBBytesPerStage, &tmap_gh, &bars->full[stage], kb * kTcBlockK,
                                         (int)(t * kTcBlockN + rank * (kTcBlockN / 2)));
                        if (++stage == p.stages) { stage = 0; phase ^= 1; }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer (leader CTA only) =====================
        if (lane == 0 && leader) {
            int stage = 0, acc = 0;
            uint32_t phase = 0, a_phase = 0, acc_phase = 0;
            for (int64_t u = pair0; u < n_units; u += pair_step) {
                const int64_t grp = u / n_qpairs;
                const int64_t t0 = p.tile_begin + grp * p.tiles_per_group;
                const int64_t t1 = (t0 + p.tiles_per_group < p.tile_end) ? t0 + p.tiles_per_group : p.tile_end;
                mbar_wait(&bars->a_full, a_phase);
                a_phase ^= 1;
                tcgen05_fence_after();
                for (int64_t t = t0; t < t1; t++) {
                    mbar_wait(&bars->tmem_empty[acc], acc_phase ^ 1);
                    tcgen05_fence_after();
                    const uint32_t tmem_d = tmem_base + (uint32_t)acc * kTcBlockN;
                    for (int kb = 0; kb < p.k_blocks; kb++) {
                        mbar_wait(&bars->full[stage], phase);
                        tcgen05_fence_after();
                        const uint64_t da = make_sw128_desc(smem_u32(smem_a + (size_t)kb * kTcABytesPerKb));
                        const uint64_t db = make_sw128_desc(smem_u32(smem_b + (size_t)stage * kTcPairBBytesPerStage));
#pragma unroll
                        for (int k4 = 0; k4 < kTcBlockK / kTcUmmaK; k4++)
                            umma_f16_pair(tmem_d, da + (uint64_t)(k4 * 2), db + (uint64_t)(k4 * 2), p.idesc, (kb | k4) != 0 ? 1u : 0u);
                        tcgen05_commit_pair(&bars->empty[stage]);
                        if (++stage == p.stages) { stage = 0; phase ^= 1; }
                    }
                    tcgen05_commit_pair(&bars->tmem_full[acc]);
                    if (++acc == 2) { acc = 0; acc_phase ^= 1; }
                }
                tcgen05_commit_pair(&bars->a_empty);
            }
        }
    } else if (warp >= 4) {
        // ===================== epilogue (both CTAs): own query tile, own TMEM lanes =====================
        const int ew = warp & 3;
        const int row = ew * 32 + lane;
        int acc = 0;
        uint32_t acc_phase = 0;
        constexpr bool QUEUED = RS >= 16;   // as in the single-CTA kernel: long lists admit through a per-lane queue
        float qv[kTcQueue];
        int qi[kTcQueue];
        int qn = 0;
#pragma unroll
        for (int r = 0; r < kTcQueue; r++) { qv[r] = -INFINITY; qi[r] = -1; }
        float rs[RS];
        int ri[RS];
        for (int64_t u = pair0; u < n_units; u += pair_step) {
            const int64_t qp = u % n_qpairs, grp = u / n_qpairs;
            const int64_t qt = 2 * qp + rank;
            const int64_t t0 = p.tile_begin + grp * p.tiles_per_group;
            const int64_t t1 = (t0 + p.tiles_per_group < p.tile_end) ? t0 + p.tiles_per_group : p.tile_end;
            const int64_t unit_n0 = t0 * kTcBlockN;
#pragma unroll
            for (int j = 0; j < RS; j++) { rs[j] = -INFINITY; ri[j] = -1; }
            const int64_t q = qt * kTcBlockM + row;
            const bool q_live = q < p.n_query;
            float kth = -INFINITY;
            float gthr = q_live ? ld_relaxed_f32(p.thr + q) : INFINITY;
            float adm = gthr;
            float published = gthr;
            float pub_share = -INFINITY;
            for (int64_t t = t0; t < t1; t++) {
                mbar_wait(&bars->tmem_full[acc], acc_phase);
                tcgen05_fence_after();
                const int64_t n0 = t * kTcBlockN;
                const int valid = (int)((p.n_gallery - n0) < kTcBlockN ? (p.n_gallery - n0) : kTcBlockN);
                const int col0 = (int)(n0 - unit_n0);
                const uint32_t taddr = tmem_base + ((uint32_t)(ew * 32) << 16) + (uint32_t)acc * kTcBlockN;
#pragma unroll 1
                for (int c0 = 0; c0 < kTcBlockN; c0 += 32) {
                    float v[32];
                    tmem_ld_32x32(taddr + (uint32_t)c0, v);
                    if (c0 + 32 > valid) {
#pragma unroll
                        for (int j = 0; j < 32; j++) v[j] = (c0 + j < valid) ? v[j] : -INFINITY;
                    }
                    tc_admit32<RS, QUEUED>(v, col0 + c0, p.k, gthr, adm, kth, rs, ri, qv, qi, qn);
                }
                tcgen05_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive_leader(&bars->tmem_empty[acc]);
                if (++acc == 2) { acc = 0; acc_phase ^= 1; }
                if (q_live) tc_exchange_bounds<RS>(p, q, grp, rs, kth, published, pub_share, gthr, adm);
            }
            if (QUEUED && __any_sync(0xffffffffu, qn > 0)) drain_queue<RS>(rs, ri, qv, qi, qn);
            if (q_live) {
                int nv = 0;
#pragma unroll
                for (int j = 0; j < RS; j++) nv += (j < p.k && ri[j] >= 0) ? 1 : 0;
                if (nv > 0) {
                    const int64_t o = q * p.cand_cap + atomicAdd(p.cand_cnt + q, nv);
#pragma unroll
                    for (int j = 0; j < RS; j++)
                        if (j < nv) {
                            p.cand_scores[o + j] = rs[j];
                            p.cand_idx[o + j] = p.idx_base + unit_n0 + ri[j];
                        }
                }
            }
        }
    }

    tcgen05_fence_before();
    cluster_sync_all();                       // neither CTA leaves (or frees TMEM) while its peer may still touch it
    if (warp == 2) {
        tcgen05_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
    }
}

// ---- host side ---------------------------------------------------------------------------------
// row-major bf16 [rows, dim] -> boxes of (64 of K) x box_rows rows, 128-byte swizzled, zero fill out of bounds
static int make_bf16_map(CUtensorMap *map, const void *base, int64_t rows, int dim, int box_rows, int box_cols = kTcBlockK,
                         CUtensorMapSwizzle swizzle = CU_TENSOR_MAP_SWIZZLE_128B,
                         CUtensorMapDataType dtype = CU_TENSOR_MAP_DATA_TYPE_BFLOAT16)
{
    // descriptors of recently used (base, shape, box) combinations are kept per thread: a serving loop calls with the
    // same gallery (and the same staging buffers) over and over, and an encode is a driver call
    struct Key {
        const void *base;
        int64_t rows;
        int dim, box_rows, box_cols, swizzle, dtype;
    };
    struct Entry {
        Key k;
        CUtensorMap m;
    };
    static thread_local Entry cache[8];
    static thread_local int cache_n = 0, cache_next = 0;
    const Key key{base, rows, dim, box_rows, box_cols, (int)swizzle, (int)dtype};
    for (int i = 0; i < cache_n; i++) {
        const Key &c = cache[i].k;
        if (c.base == key.base && c.rows == key.rows && c.dim == key.dim && c.box_rows == key.box_rows && c.box_cols == key.box_cols &&
            c.swizzle == key.swizzle && c.dtype == key.dtype) {
            *map = cache[i].m;
            return FRB_OK;
        }
    }
    EncodeTiledFn enc = get_encode_fn();
    if (!enc) {
        set_error("cuTensorMapEncodeTiled is not available from the CUDA driver");
        return FRB_ERR_CUDA;
    }
    cuuint64_t dims[2] = {(cuuint64_t)dim, (cuuint64_t)(rows > 0 ? rows : 1)};
    cuuint64_t strides[1] = {(cuuint64_t)dim * 2};
    cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(map, dtype, 2, const_cast<void *>(base), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled failed with CUresult %d (rows=%lld dim=%d)", (int)r, (long long)rows, dim);
        return FRB_ERR_CUDA;
    }
    cache[cache_next].k = key;
    cache[cache_next].m = *map;
    cache_next = (cache_next + 1) % 8;
    if (cache_n < 8) cache_n++;
    return FRB_OK;
}

// Two launches of the same kernel.  A short warm-up pass scans the first few gallery tiles so that every
// query has a realistic admission threshold before the main pass starts (otherwise all CTAs begin with
// -inf thresholds at the same moment and spend their first tiles on the slow admission path); the main
// pass covers the remaining tiles.  Both write candidate lists into one [n_groups, Q, k] array.
struct TcPass {
    int64_t tile_begin, tile_end, tiles_per_group, n_groups, group_base;
};

struct TcPlan {
    int64_t n_qtiles, n_tiles, n_groups;  // n_groups = total candidate lists per query
    TcPass warm, main;                     // warm.n_groups == 0: single pass
    size_t qbf16_bytes, thr_bytes, cnt_bytes, share_bytes, idx_bytes, score_bytes;
};

static void tc_split(int64_t tile_begin, int64_t tile_end, int64_t want_groups, int64_t group_base, TcPass *ps)
{
    const int64_t tiles = tile_end - tile_begin;
    if (want_groups > tiles) want_groups = tiles;
    if (want_groups < 1) want_groups = 1;
    ps->tile_begin = tile_begin;
    ps->tile_end = tile_end;
    ps->tiles_per_group = tiles > 0 ? (tiles + want_groups - 1) / want_groups : 1;
    ps->n_groups = tiles > 0 ? (tiles + ps->tiles_per_group - 1) / ps->tiles_per_group : 1;
    ps->group_base = group_base;
}

// Group count whose round-robin schedule (CTA p runs units p, p + P, ...) finishes earliest.  The rule of thumb above
// (16 units per CTA) is exact when n_qtiles divides P * 16 (32 query tiles: 2368 units on 148 CTAs) but loses a whole
// round of ~50-tile units otherwise: 256 query tiles x 10 groups = 2560 units = 17.3 rounds -> 18 (measured on one GPU,
// same flops: 2.74 ms at 4096 x 1M against 2.99 ms at 32768 x 125k, profiles/r2_tc_shapes.txt).  Cost model per unit:
// its tiles + ~1 tile-time (the query tile is single-buffered: the next unit's 128 KB A load starts when the last MMA
// of this one retires).  Evaluated exactly per CTA for every feasible group count.
static int64_t tc_balanced_groups(int64_t tiles, int64_t n_qt, int P, int64_t fallback)
{
    static thread_local int64_t c_tiles = -1, c_qt = -1, c_best = 0;
    static thread_local int c_P = 0;
    if (tiles == c_tiles && n_qt == c_qt && P == c_P) return c_best;
    const double ov = 1.0;
    int64_t g_max = (int64_t)32 * P / n_qt + 1;
    if (g_max > tiles) g_max = tiles;
    if (g_max > 1024) g_max = 1024;
    double best_t = 1e300;
    int64_t best = fallback;
    for (int64_t g = 1; g <= g_max; g++) {
        const int64_t tpg = (tiles + g - 1) / g;
        if ((tiles + tpg - 1) / tpg != g) continue;           // same split as a smaller g
        if (tpg < 12 && g > 1) break;                          // shorter units only add per-unit overhead
        if (tpg > 64) continue;                                // measured: units of ~120 tiles ran 8 % slower than the model says
        const int64_t U = g * n_qt, last0 = U - n_qt, last_len = tiles - (g - 1) * tpg;
        double worst = 0.0;
        for (int p = 0; p < P && p < U; p++) {
            const int64_t cnt = (U - 1 - p) / P + 1;                                             // units of CTA p
            const int64_t cl = last0 > p ? (U - 1 - p) / P - (last0 - 1 - p) / P : cnt;          // ... of which in the last group
            const double t = (double)(cnt - cl) * ((double)tpg + ov) + (double)cl * ((double)last_len + ov);
            if (t > worst) worst = t;
        }
        if (worst < best_t - 1e-9) { best_t = worst; best = g; }
    }
    c_tiles = tiles; c_qt = n_qt; c_P = P; c_best = best;
    return best;
}

// CTA pairs (cosine_tc_pair_kernel) serve every list length from two query tiles up; FRB_TC_PAIR=0 keeps one CTA per tile
static int tc_pair_max_k()
{
    const char *m = getenv("FRB_TC_PAIR_MAXK");                         // experiments: longest list the pair kernel serves
    return m ? atoi(m) : kTcMaxRegK;
}
static bool tc_use_pair(int64_t nq, int k)
{
    const char *e = getenv("FRB_TC_PAIR");
    if (e && e[0] == '0') return false;
    return nq > kTcBlockM && k <= tc_pair_max_k();
}
// Long lists (k > 8) warm a k-slot list per unit, so big batches want few long units (as the single-CTA kernel always
// plans them); small batches keep the regular planner.  Interleaved A/B, k = 16 (profiles/r2_tc_pair_long.txt):
// 4096 x 1M 3.07 -> 2.91 ms, 8192 x 500k 4.11 -> 2.99, 32768 x 125k 3.57 -> 3.24; but 1024 x 1M 1.02 -> 1.10, 256 x 1M 0.46 -> 0.55.
// k = 32 / 64: 1024 x 1M 1.42 -> 1.25 / 3.57 -> 2.45 ms with long units, 256 x 1M (k = 64) 2.01 -> 3.41 without.
static bool tc_pair_long_units(int64_t nq, int k)
{
    const char *e = getenv("FRB_TC_PAIR_LONG");                         // experiments: 0 / 1 force it
    return e ? e[0] == '1' : nq >= (k <= 16 ? 2048 : 1024);
}

static TcPlan tc_plan(int64_t nq, int64_t ng, int dim, int k)
{
    TcPlan pl;
    int sms = sm_count();
    if (sms <= 0) sms = 148;
    const bool pair = tc_use_pair(nq, k);
    pl.n_qtiles = (nq + kTcBlockM - 1) / kTcBlockM;
    const int64_t real_qtiles = pl.n_qtiles;
    if (pair) {                      // schedule in units of (query-tile pair, gallery group) over sms / 2 CTA pairs
        pl.n_qtiles = (pl.n_qtiles + 1) / 2;
        sms /= 2;
    }
    pl.n_tiles = (ng + kTcBlockN - 1) / kTcBlockN;
    if (pl.n_tiles < 1) pl.n_tiles = 1;
    // warm-up: ~1/64 of the gallery, 4..32 tiles, one wave of CTAs
    int64_t warm_tiles = 0;
    // The warm-up pass pays for itself in the single-CTA kernels; with CTA pairs it no longer does (interleaved A/B,
    // profiles/r2_tc_warm.txt: 4096 x 1M 2.85 -> 2.80 ms, 4096 x 125k 0.359 -> 0.347, 1024 x 125k 0.159 -> 0.142, 256 x 1M
    // 0.253 -> 0.234, 32768 x 125k 2.89 -> 2.88; only 4096 x 250k lost, 0.669 -> 0.695), so pair plans skip it: one launch less.
    const char *warm_env = getenv("FRB_TC_WARM");     // experiments: 0 / 1 force it off / on
    const bool want_warm = warm_env ? warm_env[0] != '0' : !pair;
    if (pl.n_tiles >= 64 && want_warm) {
        warm_tiles = pl.n_tiles / 64;
        if (warm_tiles < 4) warm_tiles = 4;
        if (warm_tiles > 32) warm_tiles = 32;
    }
    if (warm_tiles > 0) {
        int64_t wg = sms / pl.n_qtiles;
        if (wg < 1) wg = 1;
        tc_split(0, warm_tiles, wg, 0, &pl.warm);
    } else {
        pl.warm = TcPass{0, 0, 1, 0, 0};
    }
    // main: a unit's gallery group is shared (through L2) by all query tiles.  Up to 16 units per CTA for balance, but a
    // unit should span ~16 gallery tiles or more: every unit reloads its 128 KB query tile and refills the MMA /
    // TMEM pipeline, which dominated small batches when units were 2 tiles long (256 queries x 1M rows).
    const int64_t main_tiles = pl.n_tiles - warm_tiles;
    const int64_t tiles_per_cta = (main_tiles * pl.n_qtiles + sms - 1) / sms;
    int64_t units_per_cta = (tiles_per_cta + 8) / 16;
    if (units_per_cta < 1) units_per_cta = 1;
    if (units_per_cta > 16) units_per_cta = 16;
    int64_t want_groups = ((int64_t)sms * units_per_cta + pl.n_qtiles - 1) / pl.n_qtiles;
    const bool long_units = k > kTcShareMinK && (!pair || tc_pair_long_units(nq, k));
    if (long_units) {
        // long lists: every unit warms its own k-slot list (~k ln(rows / k) slow-path insertions per query and unit), so
        // few long units -- at most two per CTA, and never a third round (floor, not ceil)
        if (units_per_cta > 2) units_per_cta = 2;
        want_groups = ((int64_t)sms * units_per_cta) / pl.n_qtiles;
        if (want_groups < 1) want_groups = 1;
    }
    if (want_groups > 1024) want_groups = 1024;
    {
        const char *bal = getenv("FRB_TC_BALANCE");      // experiments: 0 keeps the rule of thumb
        if (!long_units && tiles_per_cta >= 48 && !(bal && bal[0] == '0'))
            want_groups = tc_balanced_groups(main_tiles, pl.n_qtiles, sms, want_groups);
    }
    {
        const char *ut = getenv("FRB_TC_UNIT_TILES");    // experiments: force the unit length (gallery tiles per group)
        if (ut && atoll(ut) > 0) want_groups = (main_tiles + atoll(ut) - 1) / atoll(ut);
    }
    tc_split(warm_tiles, pl.n_tiles, want_groups, pl.warm.n_groups, &pl.main);
    pl.n_groups = pl.warm.n_groups + pl.main.n_groups;
    pl.n_qtiles = real_qtiles;
    size_t n = (size_t)pl.n_groups * (size_t)nq * (size_t)k;
    pl.qbf16_bytes = align_up((size_t)pl.n_qtiles * kTcBlockM * (size_t)dim * 2, 1024);
    pl.thr_bytes = align_up((size_t)nq * sizeof(float), 256);
    pl.share_bytes = k > kTcShareMinK ? align_up((size_t)nq * 2 * kTcShareMaxGroups * sizeof(float), 256) : 0;  // one array per pass
    pl.cnt_bytes = align_up((size_t)nq * sizeof(int), 256);
    pl.idx_bytes = align_up(n * sizeof(int64_t), 256);
    pl.score_bytes = align_up(n * sizeof(float), 256);
    return pl;
}

size_t cosine_tc_workspace_bytes(int64_t n_query, int64_t n_gallery, int dim, int k)
{
    TcPlan pl = tc_plan(n_query, n_gallery, dim, k);
    return pl.qbf16_bytes + pl.thr_bytes + pl.cnt_bytes + pl.share_bytes + pl.idx_bytes + pl.score_bytes;
}

// per-query scratch reset for callers that bring their own normalised bf16 queries (no prologue launch to ride on)
__global__ void __launch_bounds__(256) tc_reset_kernel(float *__restrict__ thr, int *__restrict__ cnt, int64_t n)
{
    const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (i < n) {
        thr[i] = -INFINITY;
        cnt[i] = 0;
    }
}

// queries: fp32 rows that the prologue normalises (qnorm_mode) and rounds to bf16 — or, when queries_bf16 is given, rows
// that are ALREADY normalised bf16 (e.g. all-gathered from the ranks that normalised their own slice of the batch): the
// TMA reads them in place and no prologue runs.
// op_dtype: FRB_BF16 or FRB_F16 — the 16-bit format of the gallery and of the (normalised) queries.
int launch_cosine_tc(const float *queries, const void *queries_bf16, int64_t nq, const void *gallery_bf16, int64_t ng, int dim,
                     int qnorm_mode, int k, int64_t idx_base, float *out_scores, int64_t *out_idx, void *ws, size_t ws_bytes,
                     cudaStream_t st, int op_dtype)
{
    const CUtensorMapDataType tm_dtype = op_dtype == FRB_F16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
    (void)ws_bytes;
    if (dim % kTcBlockK != 0 || dim / kTcBlockK > kTcMaxKBlocks) {
        set_error("bf16 tensor-core path: dim=%d must be a multiple of 64 and <= 512", dim);
        return FRB_ERR_UNSUPPORTED;
    }
    int dev = 0, major = 0;
    FRB_CUDA_OK(cudaGetDevice(&dev));
    FRB_CUDA_OK(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
    if (major != 10) {
        set_error("bf16 tensor-core path needs an sm_100-class GPU (tcgen05); device is sm_%d", major * 10);
        return FRB_ERR_UNSUPPORTED;
    }
    TcPlan pl = tc_plan(nq, ng, dim, k);
    char *w = (char *)ws;
    __nv_bfloat16 *qb = (__nv_bfloat16 *)w;
    float *thr = (float *)(w + pl.qbf16_bytes);
    int *cnt = (int *)(w + pl.qbf16_bytes + pl.thr_bytes);
    float *share = (float *)(w + pl.qbf16_bytes + pl.thr_bytes + pl.cnt_bytes);
    int64_t *ci = (int64_t *)(w + pl.qbf16_bytes + pl.thr_bytes + pl.cnt_bytes + pl.share_bytes);
    float *cs = (float *)(w + pl.qbf16_bytes + pl.thr_bytes + pl.cnt_bytes + pl.share_bytes + pl.idx_bytes);

    // prologue: L2-normalise in fp32, round to bf16 (rows beyond n_query are never read: TMA zero-fills)
    // the same launch resets the shared admission thresholds (-inf) and the candidate counters (0)
    int rc = FRB_OK;
    if (queries_bf16) {
        qb = (__nv_bfloat16 *)const_cast<void *>(queries_bf16);
        tc_reset_kernel<<<(unsigned)((nq + 255) / 256), 256, 0, st>>>(thr, cnt, nq);
        FRB_LAUNCH_OK("tc_reset_kernel");
    } else {
        rc = normalize_rows_impl(queries, nq, dim, qnorm_mode, qb, op_dtype, thr, cnt, st);
        if (rc != FRB_OK) return rc;
    }
    if (pl.share_bytes) FRB_CUDA_OK(cudaMemsetAsync(share, 0xFF, pl.share_bytes, st));   // 0xFFFFFFFF = not published

    CUtensorMap tq, tg, tpf;
    rc = make_bf16_map(&tq, qb, nq, dim, kTcBlockM, kTcBlockK, CU_TENSOR_MAP_SWIZZLE_128B, tm_dtype);
    if (rc != FRB_OK) return rc;
    rc = make_bf16_map(&tg, ng > 0 ? gallery_bf16 : (const void *)qb, ng, dim, kTcBlockN, kTcBlockK, CU_TENSOR_MAP_SWIZZLE_128B, tm_dtype);
    if (rc != FRB_OK) return rc;

    // prefetch view of the gallery: unswizzled boxes of (up to) 256 of K x 256 rows, L2 only
    rc = make_bf16_map(&tpf, ng > 0 ? gallery_bf16 : (const void *)qb, ng, dim, kTcBlockN, dim < 256 ? dim : 256,
                       CU_TENSOR_MAP_SWIZZLE_NONE, tm_dtype);
    if (rc != FRB_OK) return rc;

    TcParams p;
    p.n_query = nq;
    p.n_gallery = ng;
    p.k_blocks = dim / kTcBlockK;
    p.idesc = op_dtype == FRB_F16 ? kTcIdescF16 : kTcIdescBf16;
    const size_t a_bytes = (size_t)p.k_blocks * kTcABytesPerKb;
    const bool pair = tc_use_pair(nq, k);
    int stages = pair ? (int)((kTcSmemLimit - 1024 - sizeof(TcPairBarriers) - a_bytes) / kTcPairBBytesPerStage)
                      : (int)((kTcSmemLimit - 1024 - sizeof(TcBarriers) - a_bytes) / kTcBBytesPerStage);
    if (stages > (pair ? kTcPairStagesMax : kTcMaxStages)) stages = pair ? kTcPairStagesMax : kTcMaxStages;
    if (stages < 2) {
        set_error("bf16 tensor-core path: not enough shared memory for 2 gallery stages");
        return FRB_ERR_UNSUPPORTED;
    }
    p.stages = stages;
    p.n_qtiles = pl.n_qtiles;
    p.k = k;
    const char *pf_env = getenv("FRB_TC_PREFETCH_DIST");  // tuning knob for experiments
    p.idx_base = idx_base;
    p.thr = thr;
    p.cand_cnt = cnt;
    p.cand_cap = pl.n_groups * k;
    p.cand_scores = cs;
    p.cand_idx = ci;
    const size_t smem = pair ? 1024 + a_bytes + (size_t)stages * kTcPairBBytesPerStage + sizeof(TcPairBarriers)
                             : 1024 + a_bytes + (size_t)stages * kTcBBytesPerStage + sizeof(TcBarriers);
    CUtensorMap tgh;
    if (pair) {
        // this CTA's half of a gallery tile: boxes of 128 rows
        rc = make_bf16_map(&tgh, ng > 0 ? gallery_bf16 : (const void *)qb, ng, dim, kTcBlockN / 2, kTcBlockK, CU_TENSOR_MAP_SWIZZLE_128B, tm_dtype);
        if (rc != FRB_OK) return rc;
        p.idesc = op_dtype == FRB_F16 ? kTcIdescPairF16 : kTcIdescPairBf16;
        static thread_local int pair_dev = -1;
        static thread_local size_t pair_smem = 0;
        if (pair_dev != dev || pair_smem < smem) {
            FRB_CUDA_OK(cudaFuncSetAttribute(cosine_tc_pair_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            FRB_CUDA_OK(cudaFuncSetAttribute(cosine_tc_pair_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            FRB_CUDA_OK(cudaFuncSetAttribute(cosine_tc_pair_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            FRB_CUDA_OK(cudaFuncSetAttribute(cosine_tc_pair_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            pair_dev = dev;
            pair_smem = smem;
        }
    }
    const int variant = k <= 8 ? 0 : (k <= 16 ? 1 : (k <= 32 ? 2 : (k <= kTcMaxRegK ? 3 : 4)));  // register list of 8 / 16 / 32 / 64 slots, else local memory
    typedef void (*TcKernel)(const CUtensorMap, const CUtensorMap, const CUtensorMap, const TcParams);
    const TcKernel kernels[5] = {cosine_tc_kernel<8>, cosine_tc_kernel<16>, cosine_tc_kernel<32>, cosine_tc_kernel<64>, cosine_tc_kernel<0>};
    const TcKernel kernel = kernels[variant];
    {
        // opt in to >48 KB dynamic shared memory once per (device, kernel variant, size)
        static thread_local int attr_dev[5] = {-1, -1, -1, -1, -1};
        static thread_local size_t attr_smem[5] = {0, 0, 0, 0, 0};
        if (attr_dev[variant] != dev || attr_smem[variant] < smem) {
            FRB_CUDA_OK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            attr_dev[variant] = dev;
            attr_smem[variant] = smem;
        }
    }
    const int sms = sm_count();
    const TcPass passes[2] = {pl.warm, pl.main};
    for (int ip = 0; ip < 2; ip++) {
        const TcPass &ps = passes[ip];
        if (ps.n_groups == 0) continue;
        p.tiles_per_group = ps.tiles_per_group;
        p.n_groups = ps.n_groups;
        const bool use_share = pl.share_bytes && ps.n_groups >= 2 && ps.n_groups <= kTcShareMaxGroups;
        p.share = use_share ? share + (size_t)ip * nq * kTcShareMaxGroups : nullptr;
        p.share_rank = use_share ? (int)((k + ps.n_groups - 1) / ps.n_groups) : k;
        p.tile_begin = ps.tile_begin;
        p.tile_end = ng > 0 ? ps.tile_end : ps.tile_begin;  // empty gallery: units run with no tiles and emit empty lists
        const int64_t sched_qtiles = pair ? (pl.n_qtiles + 1) / 2 : pl.n_qtiles;     // query tiles, or pairs of them
        const int64_t n_units = sched_qtiles * ps.n_groups;
        const int slots = pair ? sms / 2 : sms;
        const int grid = (int)(n_units < slots ? n_units : slots);                   // CTAs, or CTA pairs
        {
            const size_t tile_bytes = (size_t)kTcBlockN * (size_t)p.k_blocks * kTcBlockK * 2;
            const size_t groups_in_flight = (size_t)((grid + sched_qtiles - 1) / sched_qtiles);
            const size_t fit = kTcPrefetchL2Budget / (groups_in_flight * tile_bytes);
            p.prefetch_dist = pf_env ? atoi(pf_env) : (int)(fit < (size_t)kTcPrefetchDist ? fit : (size_t)kTcPrefetchDist);
        }
        ProfileScope prof(FRB_K_COSINE_TC, st);
        if (pair && k <= 8)
            cosine_tc_pair_kernel<8><<<2 * grid, kTcThreads, smem, st>>>(tq, tgh, tpf, p);
        else if (pair && k <= 16)
            cosine_tc_pair_kernel<16><<<2 * grid, kTcThreads, smem, st>>>(tq, tgh, tpf, p);
        else if (pair && k <= 32)
            cosine_tc_pair_kernel<32><<<2 * grid, kTcThreads, smem, st>>>(tq, tgh, tpf, p);
        else if (pair)
            cosine_tc_pair_kernel<64><<<2 * grid, kTcThreads, smem, st>>>(tq, tgh, tpf, p);
        else
            kernel<<<grid, kTcThreads, smem, st>>>(tq, tg, tpf, p);
        FRB_LAUNCH_OK("cosine_tc_kernel");
    }
    return topk_merge_compact(cs, ci, cnt, p.cand_cap, nq, k, /*largest=*/1, out_scores, out_idx, st);
}

}  // namespace frb
