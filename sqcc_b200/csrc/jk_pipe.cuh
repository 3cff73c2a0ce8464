// Fused J/K kernel, software-pipelined variant (default).  Included by jk_fused.cu.
//
// Same warp-task decomposition as jk_tiled_kernel (see jk_fused.cu): a warp owns (k-range of JK_KN sub-rows) x
// (64-column l-chunk, two columns per lane) and sweeps a list of NI x NJ row-blocks.  Differences:
//   * the ERI stream is prefetched STAGES-1 sub-rows ahead into a per-warp shared-memory ring with 16-byte
//     cp.async (LDGSTS.128, L2-only).  Every lane copies and later reads back only its own 16 bytes of each of
//     the NI*NJ rows, so the ring needs no barriers (cp.async.wait_group is per thread), costs no registers and
//     keeps WARPS * (STAGES-1) * (NI*NJ+1) * 512 B in flight per SM.  Lanes outside a triangular sub-row use
//     the zero-fill form, so the consumer loop is mask free.  The matching D2[k, l-chunk] density tile row
//     rides in the same ring stage (one more 16-byte copy per lane).  The pipeline runs across the row-block
//     boundaries of a task.
//   * the warp-uniform operands D[i,k], D[j,k] of a row-block are staged once per row-block in shared memory
//     and read back as broadcast 128-bit loads.
//   * the K~[i,k] / K~[j,k] lane partials are reduced across the warp through a conflict-free shared-memory
//     transpose (8 stores + 8 loads + 2 shuffles per sub-row) instead of a select-heavy shuffle butterfly.
// A TMA-only ring (cp.async.bulk per 512-byte piece + mbarrier) was measured slower: 512-byte bulk copies are
// limited by the TMA issue rate (3.6 TB/s J-only against 6.0 TB/s with this cp.async ring, profiles/).
#pragma once

namespace sqcc {

__device__ __forceinline__ uint32_t smem_u32(const void *ptr) { return (uint32_t)__cvta_generic_to_shared(ptr); }
__device__ __forceinline__ void cp_async16_zfill(uint32_t dst, const void *src, uint32_t src_bytes)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async16(uint32_t dst, const void *src)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait()
{
    asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}
// 64-bit pointer + 32-bit offset in units of 16 bytes, in ONE instruction (IMAD.WIDE.U32 with multiplier 16) instead
// of an add / add-with-carry pair (a multiplier of 1 is folded back into the pair by ptxas)
__device__ __forceinline__ const char *ptr_add_u32x16(const char *base, unsigned int off16)
{
    unsigned long long r;
    asm("mad.wide.u32 %0, %1, 16, %2;" : "=l"(r) : "r"(off16), "l"((unsigned long long)base));
    return reinterpret_cast<const char *>(r);
}
// xor-shuffle of a double without the volatile mov.b64 pack/unpack of the CUDA header (those survive as MOVs)
__device__ __forceinline__ double shfl_xor_f64(double v, int lane_mask)
{
    int lo = __double2loint(v), hi = __double2hiint(v);
    lo = __shfl_xor_sync(0xffffffffu, lo, lane_mask);
    hi = __shfl_xor_sync(0xffffffffu, hi, lane_mask);
    return __hiloint2double(hi, lo);
}
__device__ __forceinline__ unsigned long long globaltimer_ns()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
__device__ __forceinline__ void red_add_if(bool pred, double *ptr, double v)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.u32 p, %2, 0;\n\t"
        "@p red.global.add.f64 [%0], %1;\n\t}" ::"l"(ptr),
        "d"(v), "r"((uint32_t)pred)
        : "memory");
}

constexpr int JK_RED_STRIDE = 17;  // value stride of the reduction scratch (odd: conflict-free 4x4 read pattern)

// NVP = number of 8-value passes of the per-sub-row reduction: ceil(nspin_k * (NI + NJ) / 8)
template <int NI, int NJ, int STAGES, bool SLAB, bool RING, int NVP>
constexpr int jk_pipe_warp_bytes()
{
    // J~ slab + (ring[STAGES][NI*NJ+1][512 B] | D2 tile slab) + uniform-operand staging [KN][8 NVP] + reduction scratch
    return (SLAB ? JK_KN * 32 * 16 : 0) + (RING ? STAGES * (NI * NJ + 1) * 512 : JK_KN * 32 * 16) +
           JK_KN * 8 * NVP * 8 + 8 * NVP * JK_RED_STRIDE * 8 + 128;
}

// SLAB = true : J~[k, l-chunk] is accumulated over all row-blocks of a task in a warp-private shared-memory slab;
// SLAB = false: it is flushed every sub-row with red.global.add.f64 (no slab: more warps / ring stages fit).
// RING = true : ERI stream staged through the shared-memory cp.async ring (STAGES deep);
// RING = false: ERI stream loaded straight into registers with 128-bit ld.global.nc, one sub-row ahead of the FMA
//               block (no shared-memory traffic for the stream; the D2 tile is staged once per task in a slab).
template <int NI, int NJ, int NSPIN, bool WANTK, int WARPS, int STAGES, bool SLAB, bool RING>
__global__ void __launch_bounds__(WARPS * 32, 1) jk_pipe_kernel(const JKParams p)
{
    constexpr int R = NI * NJ;
    constexpr int RS = R + 1;  // ring rows per stage: R ERI rows + the D2 tile row
    constexpr int NS = WANTK ? NSPIN : 1;
    constexpr int NV = NS * (NI + NJ);  // lane-partial values reduced per sub-row
    constexpr int NVP = (NV + 7) / 8;   // reduction passes (8 values per pass)
    constexpr int WARP_BYTES = jk_pipe_warp_bytes<NI, NJ, STAGES, SLAB, RING, NVP>();
    static_assert(NVP <= 2, "reduction scratch holds 16 values");
    static_assert(NI * NJ <= 16, "s_w holds 16 values");
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned char *wbase = smem_raw + (size_t)warp * WARP_BYTES;
    double2 *s_jt = reinterpret_cast<double2 *>(wbase);              // [KN][32]
    double2 *ring = s_jt + (SLAB ? JK_KN * 32 : 0);                               // [STAGES][RS][32]
    double2 *s_d2 = ring;                                            // [KN][32] (RING == false only)
    double *s_u = reinterpret_cast<double *>(ring + (RING ? STAGES * RS * 32 : JK_KN * 32));  // [KN][8]
    double *s_red = s_u + JK_KN * 8 * NVP;                           // [8 NVP][JK_RED_STRIDE]
    double *s_w = s_red + 8 * NVP * JK_RED_STRIDE;                         // [8] D2[i_a, j_b] of the row-block
    const uint32_t ring_u32 = smem_u32(ring) + 16u * lane;
    const int n = p.n, ldd = p.ldd;
    const size_t nl = (size_t)n * ldd;
    uint32_t stage_w = 0, stage_r = 0;  // ring write / read positions, persist across tasks
    unsigned long long tr_start = 0, tr_tasks = 0;
    if (p.trace) tr_start = globaltimer_ns();

    for (;;) {
        unsigned t = 0;
        if (lane == 0) t = atomicAdd(p.counter, 1u);
        t = __shfl_sync(0xffffffffu, t, 0);
        if (t >= (unsigned)p.ntasks) break;
        const JKTask task = p.tasks[t];
        const int k0 = task.k0;
        const int cbase = task.chunk * JK_CHUNK;
        const int l0 = cbase + 2 * lane;
        const bool lane_in = l0 < ldd;
        const int kk_lo = cbase > k0 ? cbase - k0 : 0;
        const int kk_hi = (k0 + JK_KN <= n) ? JK_KN : (n - k0);
        const int nkk = kk_hi - kk_lo;
        const bool full_chunk = k0 >= cbase + JK_CHUNK - 1;
        if (nkk <= 0) continue;
        const int niter = task.rb_count * nkk;
        ++tr_tasks;

#pragma unroll 4
        for (int kk = 0; kk < (SLAB ? JK_KN : 0); ++kk) s_jt[kk * 32 + lane] = make_double2(0.0, 0.0);

        // row-block list of the task, one entry per lane (rb_count <= 32)
        int my_i0 = 0, my_j0 = 0;
        unsigned int my_mask = 0;
        if (lane < task.rb_count) {
            const RowBlock rbk = p.rowblocks[task.rb_begin + lane];
            my_i0 = rbk.i0;
            my_j0 = rbk.j0;
            my_mask = rbk.mask;
        }
        // D2 tile pointer of this lane (row k0, columns l0, l0+1); lanes past the matrix edge zero-fill
        const double *d2p = p.d2 + (size_t)k0 * ldd + (lane_in ? l0 : 0);

        // ---- producer: this lane's 16 bytes of each row of the producer's row-block, one sub-row per call
        int prod_it = 0, prod_rbi = -1, prod_kk = kk_hi;  // forces a row-block load on the first issue
        int prod_i0 = 0, prod_j0 = 0, prod_mode = 2;      // 0: unmasked, 1: one mask (l0 <= k), 2: per-row masks
        const char *pbase = reinterpret_cast<const char *>(p.P);   // running byte pointer: first valid row, sub-row k
        const char *d2row = reinterpret_cast<const char *>(d2p);   // running byte pointer into the D2 tile
        unsigned int poff[R];  // offsets of the rows relative to pbase in units of 16 bytes (rows are 128-byte aligned)
        unsigned int prv = 0;  // bit r: row r of the producer's row-block is a stored row of this shard
        double2 xn[R];         // RING == false: the next sub-row's integrals, in flight during the current FMA block
        auto issue_next = [&]() {
            if (prod_kk == kk_hi) {  // advance to the next row-block
                prod_kk = kk_lo;
                ++prod_rbi;
                prod_i0 = __shfl_sync(0xffffffffu, my_i0, prod_rbi);
                prod_j0 = __shfl_sync(0xffffffffu, my_j0, prod_rbi);
                prv = __shfl_sync(0xffffffffu, my_mask, prod_rbi);
                const bool all_valid = prv == (1u << R) - 1u;   // every row of the block is held by this shard
                int64_t first = -1;
                if (all_valid) {
                    // interior row-block (the common case): row(i,j) = ioff[i] + j Te(i) + Te(j), no per-row tests
                    int64_t tj[NJ];
#pragma unroll
                    for (int b = 0; b < NJ; ++b) tj[b] = subrow_offset(prod_j0 + b);
#pragma unroll
                    for (int a = 0; a < NI; ++a) {
                        const int64_t ti = subrow_offset(prod_i0 + a);
                        const int64_t ra = p.ioff[prod_i0 + a] + (int64_t)prod_j0 * ti;
                        if (a == 0) first = ra + tj[0];
#pragma unroll
                        for (int b = 0; b < NJ; ++b) poff[a * NJ + b] = (unsigned int)((ra + b * ti + tj[b] - first) >> 1);
                    }
                } else {
#pragma unroll
                    for (int r = 0; r < R; ++r) {
                        const int i = prod_i0 + r / NJ, j = prod_j0 + r % NJ;
                        const bool ok = (prv >> r) & 1u;
                        int64_t off = 0;
                        if (ok) {
                            off = p.ioff[i] + (int64_t)j * subrow_offset(i) + subrow_offset(j);
                            if (first < 0) first = off;
                        }
                        poff[r] = ok ? (unsigned int)((off - first) >> 1) : 0u;
                    }
                }
                pbase = reinterpret_cast<const char *>(p.P + (first < 0 ? 0 : first) + l0 + subrow_offset(k0 + kk_lo));
                d2row = reinterpret_cast<const char *>(d2p + (size_t)kk_lo * ldd);
                prod_mode = (all_valid && prod_i0 >= k0 + JK_KN) ? (full_chunk ? 0 : 1) : 2;
            }
            const int k = k0 + prod_kk;
            const uint32_t dst = ring_u32 + stage_w * (RS * 512u);
            if (!RING) {
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    bool on = true;
                    if (prod_mode == 1) on = l0 <= k;
                    if (prod_mode == 2) {
                        const int i = prod_i0 + r / NJ;
                        const int lmax = !((prv >> r) & 1u) ? -1 : (k < i ? k : (k == i ? prod_j0 + r % NJ : -1));
                        on = l0 <= lmax;
                    }
                    xn[r] = on ? ldg_stream(reinterpret_cast<const double *>(ptr_add_u32x16(pbase, poff[r]))) : make_double2(0.0, 0.0);
                }
            } else if (prod_mode == 0) {
#pragma unroll
                for (int r = 0; r < R; ++r) cp_async16(dst + r * 512u, ptr_add_u32x16(pbase, poff[r]));
                cp_async16(dst + R * 512u, d2row);
            } else if (prod_mode == 1) {
                const uint32_t nb = l0 <= k ? 16u : 0u;
#pragma unroll
                for (int r = 0; r < R; ++r) cp_async16_zfill(dst + r * 512u, ptr_add_u32x16(pbase, poff[r]), nb);
                cp_async16_zfill(dst + R * 512u, d2row, lane_in ? 16u : 0u);
            } else {
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    const int i = prod_i0 + r / NJ;
                    const int lmax = !((prv >> r) & 1u) ? -1 : (k < i ? k : (k == i ? prod_j0 + r % NJ : -1));
                    cp_async16_zfill(dst + r * 512u, ptr_add_u32x16(pbase, poff[r]), l0 <= lmax ? 16u : 0u);
                }
                cp_async16_zfill(dst + R * 512u, d2row, lane_in ? 16u : 0u);
            }
            pbase += (size_t)pad_sub(k + 1) * 8;   // next sub-row of every row
            d2row += (size_t)ldd * 8;
            stage_w = (stage_w + 1 == STAGES) ? 0 : stage_w + 1;
            ++prod_it;
            ++prod_kk;
        };
        if (RING) {
#pragma unroll
            for (int s = 0; s < STAGES - 1; ++s) {
                if (prod_it < niter) issue_next();
                cp_async_commit();
            }
        } else {
            __syncwarp();  // previous task's tile reads are done
#pragma unroll 4
            for (int kk = 0; kk < JK_KN; ++kk)
                s_d2[kk * 32 + lane] = (k0 + kk < n && lane_in) ? ldg2(p.d2 + (size_t)(k0 + kk) * ldd + l0) : make_double2(0.0, 0.0);
            __syncwarp();
#pragma unroll
            for (int r = 0; r < R; ++r) xn[r] = make_double2(0.0, 0.0);
            issue_next();
        }

        int cur_i0 = -1;
        double2 accA[NS][NI], di[NS][NI];
#pragma unroll
        for (int s = 0; s < NS; ++s)
#pragma unroll
            for (int a = 0; a < NI; ++a) accA[s][a] = di[s][a] = make_double2(0.0, 0.0);

        auto flushA = [&]() {
            if (WANTK && cur_i0 >= 0) {
#pragma unroll
                for (int s = 0; s < NS; ++s)
#pragma unroll
                    for (int a = 0; a < NI; ++a) {
                        const bool ok = lane_in && (cur_i0 + a < n);
                        double *dst = p.kt + s * nl + (size_t)(ok ? cur_i0 + a : 0) * ldd + (ok ? l0 : 0);
                        red_add_if(ok, dst, accA[s][a].x);
                        red_add_if(ok, dst + 1, accA[s][a].y);
                    }
            }
        };

        for (int rbi = 0; rbi < task.rb_count; ++rbi) {
            const int i0 = __shfl_sync(0xffffffffu, my_i0, rbi), j0 = __shfl_sync(0xffffffffu, my_j0, rbi);
            const unsigned int rmask = __shfl_sync(0xffffffffu, my_mask, rbi);
            if (i0 != cur_i0) {
                flushA();
                cur_i0 = i0;
#pragma unroll
                for (int s = 0; s < NS; ++s)
#pragma unroll
                    for (int a = 0; a < NI; ++a) {
                        accA[s][a] = make_double2(0.0, 0.0);
                        if (WANTK)
                            di[s][a] = (i0 + a < n && lane_in) ? ldg2(p.dsp + s * nl + (size_t)(i0 + a) * ldd + l0)
                                                                : make_double2(0.0, 0.0);
                    }
            }
            double2 dj[NS][NJ], accB[NS][NJ];
#pragma unroll
            for (int s = 0; s < NS; ++s)
#pragma unroll
                for (int b = 0; b < NJ; ++b) {
                    accB[s][b] = make_double2(0.0, 0.0);
                    dj[s][b] = make_double2(0.0, 0.0);
                    if (WANTK && j0 + b < n && lane_in) dj[s][b] = ldg2(p.dsp + s * nl + (size_t)(j0 + b) * ldd + l0);
                }
            double accJ[NI][NJ];
#pragma unroll
            for (int a = 0; a < NI; ++a)
#pragma unroll
                for (int b = 0; b < NJ; ++b) accJ[a][b] = 0.0;
            // stage the warp-uniform operands of this row-block in shared memory:
            //   s_w[r]      = D2[i_a, j_b] (0 for rows that are not in the shard)
            //   s_u[kk][q]  = D_s[row_q, k0+kk], q = s*(NI+NJ) + (a | NI+b); rows/k clamped into range
            //                 (out-of-range operands only ever meet zero-filled integrals)
            __syncwarp();
            if (lane < R) {
                const int i = i0 + lane / NJ, j = j0 + lane % NJ;
                const bool rvalid = (rmask >> lane) & 1u;
                s_w[lane] = rvalid ? __ldg(p.d2 + (size_t)i * ldd + j) : 0.0;
            }
            if (WANTK && lane < JK_KN) {
                // 32-bit indexing: nspin * n * ldd < 2^31 for every supported n (<= 4096)
                const int k = (k0 + lane) < n ? (k0 + lane) : n - 1;
                const double *dk = p.dsp + k;
                double *su = s_u + lane * (8 * NVP);
#pragma unroll
                for (int q = 0; q < NV; ++q) {
                    const int s = q / (NI + NJ), w = q % (NI + NJ);
                    int row = w < NI ? i0 + w : j0 + (w - NI);
                    row = row < n ? row : n - 1;
                    su[q] = __ldg(dk + (s * n + row) * ldd);
                }
            }
            __syncwarp();

            // The cross-lane reduction of the K~[.,k] partials of sub-row k is deferred by one iteration: the
            // partials are written to the transpose scratch at the end of iteration k, read back at the top of
            // iteration k+1 and summed / flushed after that iteration's FMA block has been issued.
            // Reduction scheme: one xor-16 shuffle halves the partials to 16 lanes, which store them (1 wavefront per
            // value); lane t then sums 4 entries of value t>>2 (stride 17: the 4x4 read pattern of a half-warp is
            // conflict free) and two more shuffles finish the sum.
            // (pass h handles values 8h .. 8h+7: value 8h + (lane >> 2), 4 lanes per value, 4 scratch entries each)
            const double *red_rd = s_red + (lane >> 2) * JK_RED_STRIDE + (lane & 3) * 4;
            double *red_wr = s_red + (lane & 15);
            double *red_dst[NVP];
            bool red_ok[NVP];
#pragma unroll
            for (int h = 0; h < NVP; ++h) {
                red_dst[h] = p.kt;
                red_ok[h] = false;
                if (WANTK) {
                    const int v = 8 * h + (lane >> 2);
                    const int s = v / (NI + NJ), w = v % (NI + NJ);
                    const int row = w < NI ? i0 + w : j0 + (w - NI);
                    red_ok[h] = (lane & 3) == 0 && v < NV && row < n;
                    if (red_ok[h]) red_dst[h] = p.kt + s * nl + (size_t)row * ldd;
                }
            }
            auto red_finish = [&](int k) {
#pragma unroll
                for (int h = 0; h < NVP; ++h) {
                    const double *rd = red_rd + 8 * h * JK_RED_STRIDE;
                    const bool on = 8 * h + (lane >> 2) < NV;
                    double r4[4];
#pragma unroll
                    for (int e = 0; e < 4; ++e) r4[e] = on ? rd[e] : 0.0;
                    double sum = (r4[0] + r4[1]) + (r4[2] + r4[3]);
                    sum += shfl_xor_f64(sum, 1);
                    sum += shfl_xor_f64(sum, 2);
                    red_add_if(red_ok[h], red_dst[h] + (red_ok[h] ? k : 0), sum);
                }
            };

            for (int kk = kk_lo; kk < kk_hi; ++kk) {
                const int k = k0 + kk;
                double2 xc[R];
                const double2 *buf = ring + (size_t)stage_r * (RS * 32) + lane;
                if (RING) {
                    if (prod_it < niter) issue_next();
                    cp_async_commit();
                    cp_async_wait<STAGES - 1>();  // this lane's 16 bytes of every row of the oldest stage have landed
                    stage_r = (stage_r + 1 == STAGES) ? 0 : stage_r + 1;
                } else {
#pragma unroll
                    for (int r = 0; r < R; ++r) xc[r] = xn[r];
                    if (prod_it < niter) issue_next();   // next sub-row's loads fly during this FMA block
                }

                double u[8 * NVP];
                if (WANTK) {
                    const double2 *up = reinterpret_cast<const double2 *>(s_u + kk * 8 * NVP);
#pragma unroll
                    for (int q = 0; q < (NV + 1) / 2; ++q) {
                        const double2 w = up[q];
                        u[2 * q] = w.x;
                        u[2 * q + 1] = w.y;
                    }
                }
                const double2 d2 = RING ? buf[R * 32] : s_d2[kk * 32 + lane];
                double2 jt2 = make_double2(0.0, 0.0);
                double cd[8 * NVP];
#pragma unroll
                for (int q = 0; q < 8 * NVP; ++q) cd[q] = 0.0;
#pragma unroll
                for (int a = 0; a < NI; ++a)
#pragma unroll
                    for (int b = 0; b < NJ; ++b) {
                        const double2 x = RING ? buf[(a * NJ + b) * 32] : xc[a * NJ + b];
                        accJ[a][b] = fma(x.x, d2.x, fma(x.y, d2.y, accJ[a][b]));
                        const double2 w2 = reinterpret_cast<const double2 *>(s_w)[(a * NJ + b) >> 1];
                        const double wij = ((a * NJ + b) & 1) ? w2.y : w2.x;
                        jt2.x = fma(x.x, wij, jt2.x);
                        jt2.y = fma(x.y, wij, jt2.y);
                        if (WANTK) {
#pragma unroll
                            for (int s = 0; s < NS; ++s) {
                                const double ui = u[s * (NI + NJ) + a], uj = u[s * (NI + NJ) + NI + b];
                                accA[s][a].x = fma(x.x, uj, accA[s][a].x);   // K~[i,l] += v D[j,k]
                                accA[s][a].y = fma(x.y, uj, accA[s][a].y);
                                accB[s][b].x = fma(x.x, ui, accB[s][b].x);   // K~[j,l] += v D[i,k]
                                accB[s][b].y = fma(x.y, ui, accB[s][b].y);
                                double &ca = cd[s * (NI + NJ) + a], &cb = cd[s * (NI + NJ) + NI + b];
                                ca = fma(x.x, dj[s][b].x, fma(x.y, dj[s][b].y, ca));   // K~[i,k] += v D[j,l]
                                cb = fma(x.x, di[s][a].x, fma(x.y, di[s][a].y, cb));   // K~[j,k] += v D[i,l]
                            }
                        }
                    }
                if (SLAB) {
                    double2 cur = s_jt[kk * 32 + lane];
                    cur.x += jt2.x;
                    cur.y += jt2.y;
                    s_jt[kk * 32 + lane] = cur;
                } else {
                    const bool ok = lane_in && l0 <= k;
                    double *dst = p.jt + (ok ? (size_t)k * ldd + l0 : 0);
                    red_add_if(ok, dst, jt2.x);
                    red_add_if(ok && l0 + 1 <= k, dst + 1, jt2.y);
                }
                if (WANTK) {
                    if (kk > kk_lo) red_finish(k - 1);   // previous sub-row's partials, written one iteration ago
                    // publish this sub-row's NV lane partials (conflict-free shared-memory transpose)
#pragma unroll
                    for (int q = 0; q < NV; ++q) cd[q] += shfl_xor_f64(cd[q], 16);
                    __syncwarp();
                    if (lane < 16) {
#pragma unroll
                        for (int q = 0; q < NV; ++q) red_wr[q * JK_RED_STRIDE] = cd[q];
                    }
                    __syncwarp();
                }
            }
            if (WANTK) {   // last sub-row of the sweep
                red_finish(k0 + kk_hi - 1);
            }

            if (WANTK) {
#pragma unroll
                for (int s = 0; s < NS; ++s)
#pragma unroll
                    for (int b = 0; b < NJ; ++b) {
                        const bool ok = lane_in && (j0 + b < n);
                        double *dst = p.kt + s * nl + (size_t)(ok ? j0 + b : 0) * ldd + (ok ? l0 : 0);
                        red_add_if(ok, dst, accB[s][b].x);
                        red_add_if(ok, dst + 1, accB[s][b].y);
                    }
            }
            {
                // J~[i_a, j_b]: reduce the R row dots across the warp (same shared-memory transpose)
#pragma unroll
                for (int a = 0; a < NI; ++a)
#pragma unroll
                    for (int b = 0; b < NJ; ++b) accJ[a][b] += shfl_xor_f64(accJ[a][b], 16);
#pragma unroll
                for (int h = 0; h < (R + 7) / 8; ++h) {   // eight row dots per pass through the scratch
                    __syncwarp();
                    if (lane < 16) {
#pragma unroll
                        for (int q = 0; q < 8; ++q)
                            if (8 * h + q < R) red_wr[q * JK_RED_STRIDE] = accJ[(8 * h + q) / NJ][(8 * h + q) % NJ];
                    }
                    __syncwarp();
                    const int v = 8 * h + (lane >> 2);
                    const double *rd = s_red + (lane >> 2) * JK_RED_STRIDE + (lane & 3) * 4;
                    double sum = 0.0;
                    if (v < R) {
#pragma unroll
                        for (int e = 0; e < 4; ++e) sum += rd[e];
                    }
                    sum += shfl_xor_f64(sum, 1);
                    sum += shfl_xor_f64(sum, 2);
                    const int i = i0 + v / NJ, j = j0 + v % NJ;
                    const bool ok = (lane & 3) == 0 && v < R && i < n && j <= i;  // rows outside the shard add 0
                    red_add_if(ok, p.jt + (ok ? (size_t)i * ldd + j : 0), sum);
                }
            }
        }
        flushA();
        if (SLAB && lane_in) {
            for (int kk = kk_lo; kk < kk_hi; ++kk) {
                const int k = k0 + kk;
                const double2 cur = s_jt[kk * 32 + lane];
                if (l0 <= k) red_add(p.jt + (size_t)k * ldd + l0, cur.x);
                if (l0 + 1 <= k) red_add(p.jt + (size_t)k * ldd + l0 + 1, cur.y);
            }
        }
    }
    if (RING) cp_async_wait<0>();
    if (p.trace && lane == 0) {
        unsigned long long *tr = p.trace + 3 * (size_t)(blockIdx.x * WARPS + warp);
        tr[0] = tr_start;
        tr[1] = globaltimer_ns();
        tr[2] = tr_tasks;
    }
}

}  // namespace sqcc
