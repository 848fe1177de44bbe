/*
 * poa_fill.cuh -- the DP fill: banded / full Needleman-Wunsch of one read against the DAG, one warp
 * per window, hand-written for sm_100a.
 *
 * Replaces vendor/GenomeWorks/cudapoa/src/cudapoa_nw.cuh:150-473 and cudapoa_nw_banded.cuh:172-482
 * (fill part) with spoa's recurrence (vendor/spoa/src/sisd_alignment_engine.cpp:283-338).
 *
 * Mapping
 *   - a row is processed in chunks of 256 columns: 32 lanes x 8 int16 cells, one 128-bit vector
 *     per lane (the reference uses 4 cells per lane and 64-bit loads, cudapoa_nw.cuh:82-99);
 *   - cells are packed two per 32-bit register and updated with the native packed instructions
 *     VIADDMNMX.S16x2 (__viaddmax_s16x2: max(a+b, c), one issue slot per two cells and
 *     predecessor term) and VIMNMX3.S16x2;
 *   - the matrix is kept in the skewed domain S = H - j*gap, which turns the horizontal dependency
 *     H[i][j-1]+gap into a plain prefix max: 14 packed ops inside the lane + a 5-step warp scan,
 *     instead of the reference's iterate-until-stable loop (cudapoa_nw.cuh:272-317);
 *   - the graph is consumed as a stream of packed 32-bit row records: 32 rows are fetched by one
 *     coalesced load (one record per lane), double-buffered, and handed out by a shuffle, so the
 *     row loop has no dependent global load for its own metadata;
 *   - the last R score rows live in a shared-memory ring (predecessors are almost always within a
 *     few ranks); only older predecessors are re-read from global/L2.  Every
 *     row is also written once to HBM for the traceback (that write is the algorithmic traffic);
 *   - the band is snapped to multiples of 8 columns, so a predecessor row with a different band
 *     start is the same 128-bit load at a lane-shifted address; cells outside a band read NEG;
 *   - match/mismatch terms come from a per-read profile in shared memory (one LDS.128 per lane per
 *     row), built once per read for A,C,G,T,N (+ one on-demand slot for any other letter).
 *
 * Must produce exactly the matrix the scalar twin in tests/emu/emu_poa.cpp (ScalarFill) produces.
 */
#pragma once
#include "poa_core.cuh"

namespace b200poa {

#if POA_DEVICE

constexpr int PROF_ROWS = 6; /* A C G T N + dynamic */

__device__ __forceinline__ uint32_t pack2(int lo, int hi) {
    return ((uint32_t)lo & 0xFFFFu) | ((uint32_t)hi << 16);
}

struct CudaFill {
    int16_t* prof;       /* shared: PROF_ROWS rows of prof_stride int16 */
    int16_t* ring;       /* shared: ring_rows rows of ring_stride int16 */
    int32_t prof_stride;
    int32_t ring_stride; /* cells per ring row (>= widest band of the batch) */
    int32_t ring_mask;   /* ring_rows - 1 (power of two) */
    int32_t dyn_code;    /* letter currently held by profile row 5, or -1 */

    __device__ __forceinline__ void build_prof_row(int row, int c, const ReadGeom& g,
                                                   const uint8_t* read, int mg, int xg) {
        const int lane = threadIdx.x & 31;
        int16_t* dst = prof + row * prof_stride;
        for (int col = lane; col < g.colsP; col += 32) {
            int v = xg;
            if (col >= 1 && col <= g.len && (int)read[col - 1] == c) v = mg;
            dst[col] = (int16_t)v;
        }
    }

    __device__ int32_t operator()(const Slot& s, const Params& p, WinState& st, const ReadGeom& g,
                                  const uint8_t* read) {
        const int lane = threadIdx.x & 31;
        const int N = st.n_nodes;
        const int mg = p.match - p.gap, xg = p.mismatch - p.gap;
        const uint32_t NEG2 = pack2(NEG, NEG);
        const uint32_t G2 = pack2(p.gap, p.gap);
        const int bw = g.bw;
        const int nchunks = (bw + CHUNK - 1) / CHUNK;
        const size_t stride = (size_t)p.stride;
        int16_t* const S = s.S;
        const uint32_t* const row_rec = s.row_rec;
        const uint32_t* const row_pred = s.row_pred;
        const int R = ring_mask + 1;

        __syncwarp();
        {   /* all five fixed profile rows in one pass over the read */
            for (int col = lane; col < g.colsP; col += 32) {
                const int ch = (col >= 1 && col <= g.len) ? (int)read[col - 1] : -1;
                prof[0 * prof_stride + col] = (int16_t)(ch == 'A' ? mg : xg);
                prof[1 * prof_stride + col] = (int16_t)(ch == 'C' ? mg : xg);
                prof[2 * prof_stride + col] = (int16_t)(ch == 'G' ? mg : xg);
                prof[3 * prof_stride + col] = (int16_t)(ch == 'T' ? mg : xg);
                prof[4 * prof_stride + col] = (int16_t)(ch == 'N' ? mg : xg);
            }
        }
        dyn_code = -1;
        /* row 0: H[0][j] = j*gap  =>  S = 0 (global copy for the traceback, ring slot 0 for the fill) */
        for (int o = lane * 8; o < bw; o += CHUNK) {
            *reinterpret_cast<uint4*>(S + o) = make_uint4(0u, 0u, 0u, 0u);
            *reinterpret_cast<uint4*>(ring + o) = make_uint4(0u, 0u, 0u, 0u);
        }
        /* row records: lane l holds the record of row (block*32 + l); next block prefetched */
        uint32_t recA = row_rec[lane];
        uint32_t recB = row_rec[32 + lane];
        __syncwarp();

        int best = NEG, end_row = 0;
        int po = 0;      /* running offset into row_pred (CSR is contiguous in row order) */
        int bs_prev = 0; /* band start of row i-1 */
        for (int i = 1; i <= N; ++i) {
            if ((i & 31) == 0) {
                recA = recB;
                recB = row_rec[i + 32 + lane];
            }
            const uint32_t rec = __shfl_sync(0xffffffffu, recA, i & 31);
            const int np = rec_npred(rec);
            const int bs = rec_bs(rec);
            int prow = rec_prow(rec);
            if (prow == 5) {
                const int code = rec_code(rec);
                if (dyn_code != code) {
                    __syncwarp();
                    build_prof_row(5, code, g, read, mg, xg);
                    dyn_code = code;
                    __syncwarp();
                }
            }
            const int16_t* profrow = prof + prow * prof_stride;
            int16_t* Srow = S + (size_t)i * stride;
            int16_t* Rrow = ring + (i & ring_mask) * ring_stride;
            uint32_t carry = NEG2; /* S[i][last column of the previous chunk], both halves */

            for (int k = 0; k < nchunks; ++k) {
                const int o0 = k * CHUNK + lane * 8; /* offset of this lane's cells in the row */
                const bool active = o0 < bw;
                const int c0 = bs + o0;              /* first column of this lane */
                uint4 P = make_uint4(0u, 0u, 0u, 0u);
                if (active) P = *reinterpret_cast<const uint4*>(profrow + c0);
                uint32_t a0 = NEG2, a1 = NEG2, a2 = NEG2, a3 = NEG2;

                for (int q = 0; q < np; ++q) {
                    int pr, bsp;
                    if (q == 0 && rec_p0prev(rec)) {
                        pr = i - 1;
                        bsp = bs_prev;
                    } else {
                        const uint32_t pe = row_pred[po + q];
                        pr = (int)(pe & 0xFFFFu);
                        bsp = (int)(pe >> 16);
                    }
                    const int off = c0 - bsp; /* offset of column c0 in the predecessor row */
                    const bool inband = active && off >= 0 && off + 8 <= bw;
                    const int lo = off - 1;   /* cell (pr, c0-1), needed by lane 0 only */
                    const bool need_left = (lane == 0) && c0 >= 1 && lo >= 0 && lo < bw;
                    uint4 V = make_uint4(NEG2, NEG2, NEG2, NEG2);
                    int lv = NEG;
                    if (i - pr < R) { /* recent row: shared-memory ring */
                        const int16_t* src = ring + (pr & ring_mask) * ring_stride;
                        if (inband) V = *reinterpret_cast<const uint4*>(src + off);
                        if (need_left) lv = src[lo];
                    } else {          /* old row: global / L2 */
                        const int16_t* src = S + (size_t)pr * stride;
                        if (inband) V = *reinterpret_cast<const uint4*>(src + off);
                        if (need_left) lv = src[lo];
                    }
                    uint32_t leftw = __shfl_up_sync(0xffffffffu, V.w, 1);
                    if (lane == 0) leftw = ((uint32_t)lv) << 16;
                    const uint32_t d0 = __funnelshift_l(leftw, V.x, 16);
                    const uint32_t d1 = __funnelshift_l(V.x, V.y, 16);
                    const uint32_t d2 = __funnelshift_l(V.y, V.z, 16);
                    const uint32_t d3 = __funnelshift_l(V.z, V.w, 16);
                    a0 = __viaddmax_s16x2(d0, P.x, a0);
                    a1 = __viaddmax_s16x2(d1, P.y, a1);
                    a2 = __viaddmax_s16x2(d2, P.z, a2);
                    a3 = __viaddmax_s16x2(d3, P.w, a3);
                    a0 = __viaddmax_s16x2(V.x, G2, a0);
                    a1 = __viaddmax_s16x2(V.y, G2, a1);
                    a2 = __viaddmax_s16x2(V.z, G2, a2);
                    a3 = __viaddmax_s16x2(V.w, G2, a3);
                }

                /* horizontal: inclusive prefix max over the 8 cells of the lane ... */
                a0 = __vmaxs2(a0, __byte_perm(a0, NEG2, 0x1054));
                a1 = __vmaxs2(a1, __byte_perm(a1, NEG2, 0x1054));
                a2 = __vmaxs2(a2, __byte_perm(a2, NEG2, 0x1054));
                a3 = __vmaxs2(a3, __byte_perm(a3, NEG2, 0x1054));
                a1 = __vmaxs2(a1, __byte_perm(a0, a0, 0x3232));
                a2 = __vmaxs2(a2, __byte_perm(a1, a1, 0x3232));
                a3 = __vmaxs2(a3, __byte_perm(a2, a2, 0x3232));
                /* ... then across lanes (packed, both halves equal) */
                uint32_t tt = __byte_perm(a3, a3, 0x3232);
                if (lane == 0) tt = __vmaxs2(tt, carry);
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) {
                    const uint32_t u = __shfl_up_sync(0xffffffffu, tt, d);
                    if (lane >= d) tt = __vmaxs2(tt, u);
                }
                uint32_t excl = __shfl_up_sync(0xffffffffu, tt, 1);
                if (lane == 0) excl = carry;
                a0 = __vimax3_s16x2(a0, excl, NEG2);
                a1 = __vimax3_s16x2(a1, excl, NEG2);
                a2 = __vimax3_s16x2(a2, excl, NEG2);
                a3 = __vimax3_s16x2(a3, excl, NEG2);
                if (nchunks > 1) {
                    carry = __shfl_sync(0xffffffffu, tt, 31);
                    carry = __vmaxs2(carry, NEG2);
                }

                if (active) {
                    const uint4 out = make_uint4(a0, a1, a2, a3);
                    *reinterpret_cast<uint4*>(Srow + o0) = out;
                    *reinterpret_cast<uint4*>(Rrow + o0) = out;
                }

                if (rec_sink(rec)) { /* sink row: candidate end cell at column len */
                    const int eo = g.len - bs - k * CHUNK;
                    if (eo >= 0 && eo < CHUNK && eo + k * CHUNK < bw) {
                        const int e = eo & 7;
                        uint32_t w = (e < 2) ? a0 : (e < 4) ? a1 : (e < 6) ? a2 : a3;
                        int val = (e & 1) ? ((int)w >> 16) : (int)(int16_t)(w & 0xFFFFu);
                        val = __shfl_sync(0xffffffffu, val, eo >> 3);
                        if (val > best) {
                            best = val;
                            end_row = i;
                        }
                    }
                }
            }
            po += np;
            bs_prev = bs;
            __syncwarp(); /* row i (ring + global) is visible to every lane before it is read */
        }
        return end_row;
    }
};

#endif /* POA_DEVICE */

} // namespace b200poa
