// spmv.cuh — the PDLP hot kernel: a persistent, warp-specialised CSR-stream SpMV with a fused epilogue.
//
//   out_r = epilogue( sum_k val[k] * vec[idx[k]],  k in [ptr[r], ptr[r+1]) )
//
// The same kernel serves A.x (CSR of A) and A'.y (CSC of A = CSR of A').  It is HBM-bound: per row it
// must move 12 B per entry (val 8 + idx 4), 4 B of row pointer, the epilogue's operand vectors and its
// outputs; the gathered vector (8n or 8m bytes) lives in the 126 MB L2.
//
// Design (sm_100a):
//   * a CTA is CW consumer warps + 1 producer warp and walks tiles of RT = 32*CW/L rows (persistent grid,
//     static round-robin);
//   * EVERYTHING that streams — the tile's slice of val/idx, its row pointers and the epilogue's operand
//     vectors (x, c, l, u, x0 / y, lc, uc, y0) — is brought into shared memory by 1-D TMA bulk copies
//     (cp.async.bulk -> mbarrier complete_tx) issued by one elected producer thread into a ring of NST
//     stages; full/empty mbarriers hand stages back and forth, there is no CTA-wide barrier in the loop,
//     so several stages of HBM traffic per CTA stay in flight while the consumers work;
//   * the only loads the consumers issue to global memory are the gathers vec[idx[k]] (random 8-byte reads
//     that hit L2, marked evict-last; the matrix stream is marked evict-first), U of them in flight per
//     thread; a group of L lanes walks one row (L = 1: one thread per row, the sum is formed strictly in
//     index order with separate multiply and add, i.e. bit-identical to a scalar CPU loop);
//   * rows longer than a stage are walked in pieces with a running sum.
// Array contract: val/idx are over-allocated by SPMV_PAD entries, ptr by 4 ints and every epilogue
// operand by SPMV_VPAD doubles (16-byte-granular copies must stay inside the allocations).
#pragma once
#include "common.cuh"
#include "tma.cuh"
#include <algorithm>
#include <cstdlib>

namespace elp {

constexpr int SPMV_PAD = 8;        // extra entries behind val / idx
constexpr int SPMV_PTR_PAD = 4;    // extra ints behind ptr
constexpr int SPMV_VPAD = 2;       // extra doubles behind every epilogue operand vector
constexpr int SPMV_MAX_NST = 8;

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

struct SpmvStageLayout {          // byte offsets inside one stage
    int cap;                      // entries per stage (multiple of 4)
    int off_idx, off_ptr, off_ops, bytes;
};
__host__ __device__ inline SpmvStageLayout spmv_stage_layout(int cap, int rt, int nin) {
    SpmvStageLayout s;
    s.cap = cap;
    s.off_idx = cap * 8;
    s.off_ptr = s.off_idx + cap * 4;
    s.off_ops = s.off_ptr + (rt + 4) * 4;
    s.bytes = s.off_ops + nin * rt * 8;
    s.bytes = (s.bytes + 127) & ~127;
    return s;
}

template <int CW, int L, class Epi>
__global__ void __launch_bounds__(CW * 32 + 32)
spmv_stream_kernel(int nrows, int ntiles, int cap, int nst, int hints, const int* __restrict__ ptr,
                   const int* __restrict__ idx, const double* __restrict__ val, const double* __restrict__ vec, Epi epi) {
    constexpr int RT = CW * 32 / L;          // rows per tile
    constexpr int NIN = Epi::NIN;
    constexpr int U = (L == 1) ? 8 : 4;      // gathers in flight per thread
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const SpmvStageLayout lay = spmv_stage_layout(cap, RT, NIN);
    unsigned char* stages = smem_raw;
    uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw + (size_t)nst * lay.bytes);   // [nst]
    uint64_t* empty = full + SPMV_MAX_NST;                                               // [nst]
    int* sinfo = reinterpret_cast<int*>(empty + SPMV_MAX_NST);                           // [nst][4]
    const int tid = threadIdx.x;
    const int warp = tid >> 5;

    if (tid == 0) {
        for (int s = 0; s < nst; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], CW); }
        mbar_fence_init();
    }
    __syncthreads();

    if (warp == CW) {
        // ---------------- producer: one elected thread feeds the ring -----------------------------
        if (tid != CW * 32) return;
        const uint64_t pol_stream = l2_policy_evict_first();
        int it = 0;
        int tile = blockIdx.x;
        int s0 = 0, e1 = 0, rows = 0;
        auto bounds = [&](int t, int& a, int& b, int& nr) {
            if (t < ntiles) {
                const int r0 = t * RT;
                nr = min(RT, nrows - r0);
                a = __ldg(ptr + r0);
                b = __ldg(ptr + r0 + nr);
            } else { a = 0; b = 0; nr = 0; }
        };
        bounds(tile, s0, e1, rows);
        while (tile < ntiles) {
            int ns0, ne1, nrows_next;
            bounds(tile + gridDim.x, ns0, ne1, nrows_next);       // in flight while this tile is issued
            const int a0 = s0 & ~3, a1 = (e1 + 3) & ~3;
            const int r0 = tile * RT;
            int piece = 0;
            for (;;) {
                const int stage = it % nst;
                if (it >= nst) mbar_wait(&empty[stage], (uint32_t)((it / nst) - 1) & 1u);
                const int pstart = a0 + piece * cap;
                const int pcnt = max(0, min(cap, a1 - pstart));
                const bool last = pstart + pcnt >= a1;
                unsigned char* sb = stages + (size_t)stage * lay.bytes;
                int* info = sinfo + stage * 4;
                info[0] = pstart;
                info[1] = max(s0, pstart) - pstart;               // first real entry of the piece
                info[2] = min(e1, pstart + pcnt) - pstart;        // one past its last real entry
                info[3] = (piece == 0 ? 1 : 0) | (last ? 2 : 0);
                const uint32_t ptr_bytes = (uint32_t)((rows + 1 + 3) & ~3) * 4u;
                const uint32_t op_bytes = (uint32_t)((rows + 1) & ~1) * 8u;
                uint32_t tx = (uint32_t)pcnt * 12u;
                if (piece == 0) tx += ptr_bytes + (uint32_t)NIN * op_bytes;
                mbar_expect_tx(&full[stage], tx);
                if (pcnt > 0) {
                    if (hints & 1) {
                        tma_load_1d_hint(sb, val + pstart, (uint32_t)pcnt * 8u, &full[stage], pol_stream);
                        tma_load_1d_hint(sb + lay.off_idx, idx + pstart, (uint32_t)pcnt * 4u, &full[stage], pol_stream);
                    } else {
                        tma_load_1d(sb, val + pstart, (uint32_t)pcnt * 8u, &full[stage]);
                        tma_load_1d(sb + lay.off_idx, idx + pstart, (uint32_t)pcnt * 4u, &full[stage]);
                    }
                }
                if (piece == 0) {
                    tma_load_1d(sb + lay.off_ptr, ptr + r0, ptr_bytes, &full[stage]);
#pragma unroll
                    for (int i = 0; i < NIN; ++i)
                        tma_load_1d(sb + lay.off_ops + i * RT * 8, epi.in(i) + r0, op_bytes, &full[stage]);
                }
                ++it;
                ++piece;
                if (last) break;
            }
            tile += gridDim.x;
            s0 = ns0; e1 = ne1; rows = nrows_next;
        }
        return;
    }

    // ---------------- consumers ------------------------------------------------------------------
    const uint64_t pol_keep = l2_policy_evict_last();
    const int g = tid / L;                 // row of the tile this lane group owns
    const int sub = tid % L;
    int q = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int row = tile * RT + g;
        int st = 0, en = 0;
        typename Epi::Pre pre{};
        double acc = 0.0;
        for (;;) {
            const int stage = q % nst;
            mbar_wait(&full[stage], (uint32_t)(q / nst) & 1u);
            const unsigned char* sb = stages + (size_t)stage * lay.bytes;
            const int* info = sinfo + stage * 4;
            const int pstart = info[0], lo = info[1], hi = info[2], flags = info[3];
            const double* sv = reinterpret_cast<const double*>(sb);
            const int* si = reinterpret_cast<const int*>(sb + lay.off_idx);
            if ((flags & 1) && row < nrows) {
                const int* sp = reinterpret_cast<const int*>(sb + lay.off_ptr);
                st = sp[g];
                en = sp[g + 1];
                if (sub == 0) pre = epi.preload(reinterpret_cast<const double*>(sb + lay.off_ops), RT, g);
            }
            const int b = max(st, pstart + lo) - pstart, f = min(en, pstart + hi) - pstart;
            for (int k0 = b + sub; k0 < f; k0 += U * L) {
                double v[U], x[U];
                int c[U];
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const int k = k0 + u * L;
                    if (k < f) { v[u] = sv[k]; c[u] = si[k]; }
                }
#pragma unroll
                for (int u = 0; u < U; ++u)
                    if (k0 + u * L < f) x[u] = (hints & 2) ? ldg_hint(vec + c[u], pol_keep) : __ldg(vec + c[u]);
#pragma unroll
                for (int u = 0; u < U; ++u)
                    if (k0 + u * L < f) acc = __dadd_rn(acc, __dmul_rn(v[u], x[u]));
            }
            __syncwarp();
            if ((tid & 31) == 0) mbar_arrive(&empty[stage]);      // this warp is done reading the stage
            ++q;
            if (flags & 2) break;
        }
        if (L > 1) acc = group_sum<L>(acc);
        if (sub == 0 && row < nrows) epi.apply(row, acc, pre);
    }
}

// ---- launch plan ---------------------------------------------------------------------------------
struct SpmvPlan {
    int cw = 8, L = 1, cap = 2048, nst = 3, ctas_per_sm = 2, ntiles = 0, hints = 3;
    int rt() const { return cw * 32 / L; }
};

inline int env_int(const char* name, int dflt) {
    const char* e = getenv(name);
    return e ? atoi(e) : dflt;
}

// lanes per row from the mean row length: short rows get one thread each, long rows a whole warp
inline int pick_lanes(int64_t nnz, int64_t nrows) {
    const double avg = nrows > 0 ? (double)nnz / (double)nrows : 0.0;
    int L = 1;
    while (L < 32 && avg >= 24.0 * L) L *= 2;      // avg < 24 -> 1, < 48 -> 2, ..., >= 384 -> 32
    return L;
}

inline size_t spmv_smem_bytes(const SpmvPlan& p, int nin) {
    return (size_t)p.nst * spmv_stage_layout(p.cap, p.rt(), nin).bytes + 2 * SPMV_MAX_NST * 8 + SPMV_MAX_NST * 16;
}

// nin_max: the largest operand count among the epilogues that will run with this plan
inline SpmvPlan plan_spmv(int64_t nnz, int nrows, int nin_max, int force_lanes = 0) {
    SpmvPlan p;
    const double avg = nrows > 0 ? (double)nnz / (double)nrows : 0.0;
    p.L = pick_lanes(nnz, nrows);
    if (force_lanes > 0) p.L = force_lanes;
    if (const int l = env_int("ELP_SPMV_L", 0)) p.L = l;       // debugging / sweeps
    if (p.L < 1 || p.L > 32 || (p.L & (p.L - 1))) p.L = 1;
    p.cw = (p.L == 1) ? env_int("ELP_SPMV_CW", 8) : 8;
    if (p.cw != 4 && p.cw != 8) p.cw = 8;
    const int rt = p.rt();
    const double mul = env_int("ELP_SPMV_CAPMUL_PCT", 125) / 100.0;
    int64_t cap = (int64_t)(avg * rt * mul) + 64;
    cap = (cap + 255) / 256 * 256;
    cap = std::max<int64_t>(256, std::min<int64_t>(cap, 8192));
    if (const int c = env_int("ELP_SPMV_CAP", 0)) cap = std::max(64, c / 4 * 4);
    p.cap = (int)cap;
    p.hints = env_int("ELP_SPMV_HINTS", 3);
    p.ntiles = ceil_div(nrows, rt);
    // ring depth and residency: fill ~200 KB of shared memory per SM with >= 3 stages per CTA
    const int stage = spmv_stage_layout(p.cap, rt, nin_max).bytes;
    const int budget = 200 * 1024;
    p.nst = env_int("ELP_SPMV_NST", 0);
    p.ctas_per_sm = env_int("ELP_SPMV_CTAS", 0);
    if (p.ctas_per_sm <= 0) {
        const int want = p.cw == 8 ? 2 : 4;
        p.ctas_per_sm = std::max(1, std::min(want, budget / (2 * stage + 256)));   // two resident CTAs beat a deeper ring
    }
    if (p.nst <= 0) p.nst = std::max(2, std::min(SPMV_MAX_NST, (budget / p.ctas_per_sm - 256) / stage));
    p.nst = std::max(2, std::min(p.nst, SPMV_MAX_NST));
    return p;
}

template <int CW, int L, class Epi>
void launch_spmv_inst(const SpmvPlan& p, int nrows, const int* ptr, const int* idx, const double* val,
                      const double* vec, const Epi& epi, cudaStream_t st) {
    auto kern = spmv_stream_kernel<CW, L, Epi>;
    static bool configured[16] = {};
    int dev = 0;
    ELP_CUDA(cudaGetDevice(&dev));
    if (dev < 16 && !configured[dev]) {
        ELP_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        configured[dev] = true;
    }
    const size_t smem = spmv_smem_bytes(p, Epi::NIN);
    ELP_REQUIRE(smem <= 227 * 1024, "spmv: stage ring of %zu bytes does not fit in shared memory", smem);
    const int grid = std::max(1, std::min(p.ntiles, kNumSMs * p.ctas_per_sm));
    ELP_LAUNCH(kern, grid, CW * 32 + 32, smem, st, nrows, p.ntiles, p.cap, p.nst, p.hints, ptr, idx, val, vec, epi);
}

template <class Epi>
void launch_spmv(const SpmvPlan& p, int nrows, const int* ptr, const int* idx, const double* val, const double* vec,
                 const Epi& epi, cudaStream_t st) {
    if (nrows <= 0) return;
    switch (p.L) {
        case 1:
            if (p.cw == 4) launch_spmv_inst<4, 1, Epi>(p, nrows, ptr, idx, val, vec, epi, st);
            else launch_spmv_inst<8, 1, Epi>(p, nrows, ptr, idx, val, vec, epi, st);
            break;
        case 2:  launch_spmv_inst<8, 2, Epi>(p, nrows, ptr, idx, val, vec, epi, st); break;
        case 4:  launch_spmv_inst<8, 4, Epi>(p, nrows, ptr, idx, val, vec, epi, st); break;
        case 8:  launch_spmv_inst<8, 8, Epi>(p, nrows, ptr, idx, val, vec, epi, st); break;
        case 16: launch_spmv_inst<8, 16, Epi>(p, nrows, ptr, idx, val, vec, epi, st); break;
        default: launch_spmv_inst<8, 32, Epi>(p, nrows, ptr, idx, val, vec, epi, st); break;
    }
}

}  // namespace elp
