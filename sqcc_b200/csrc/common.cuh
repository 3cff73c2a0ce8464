// Shared declarations for libsqcc_b200.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <string>
#include <vector>

#include "../../include/sqcc_b200.h"

namespace sqcc {

// ---------------------------------------------------------------- error plumbing
void set_error(const std::string &msg);
#define SQCC_CUDA(call)                                                                              \
    do {                                                                                             \
        cudaError_t e_ = (call);                                                                     \
        if (e_ != cudaSuccess) {                                                                     \
            sqcc::set_error(std::string(#call) + ": " + cudaGetErrorString(e_));                     \
            return SQCC_ERR_CUDA;                                                                    \
        }                                                                                            \
    } while (0)
#define SQCC_REQUIRE(cond, msg)                                                                      \
    do {                                                                                             \
        if (!(cond)) {                                                                               \
            sqcc::set_error(std::string(msg));                                                       \
            return SQCC_ERR_INVALID;                                                                 \
        }                                                                                            \
    } while (0)
#define SQCC_TRY(expr)                                                                               \
    do {                                                                                             \
        int s_ = (expr);                                                                             \
        if (s_ != SQCC_OK) return s_;                                                                \
    } while (0)

// ---------------------------------------------------------------- native v1 layout arithmetic
__host__ __device__ __forceinline__ int64_t pair_index(int64_t i, int64_t j) { return i * (i + 1) / 2 + j; }
// Every sub-row is padded to a multiple of SUBROW_ALIGN doubles (128 bytes): a warp's 512-byte piece of a
// sub-row then starts on a 128-byte line, so each 8-lane quarter of a 128-bit LDGSTS/LDG hits exactly one L1
// line (4 shared-memory fill wavefronts and 16 L2 sectors per request instead of ~10 and ~20 when sub-rows are
// only 16-byte aligned; measured with ncu, profiles/).
constexpr int SUBROW_ALIGN = 16;
__host__ __device__ __forceinline__ int64_t pad_sub(int64_t x) { return (x + (SUBROW_ALIGN - 1)) & ~(int64_t)(SUBROW_ALIGN - 1); }
// Te(k) = sum_{k' < k} pad_sub(k' + 1) = A (M + 1) (A M / 2 + r),  A = SUBROW_ALIGN, M = k / A, r = k % A.
__host__ __device__ __forceinline__ int64_t subrow_offset(int64_t k)
{
    const int64_t M = k / SUBROW_ALIGN, r = k % SUBROW_ALIGN;
    return SUBROW_ALIGN * (M + 1) * ((SUBROW_ALIGN / 2) * M + r);
}
__host__ __device__ __forceinline__ int64_t row_length(int64_t i, int64_t j) { return subrow_offset(i) + pad_sub(j + 1); }
// Offset of row (i, 0) from the start of the full tensor: sum over i' < i of [(i'+1) Te(i') + Te(i'+1)].
int64_t iblock_offset(int64_t i);   // host, O(i) loop (cached table in the context for device use)
__host__ __device__ __forceinline__ double degeneracy(int i, int j, int k, int l)
{
    double f = 1.0;
    if (i == j) f *= 0.5;
    if (k == l) f *= 0.5;
    if (i == k && j == l) f *= 0.5;
    return f;
}

// ---------------------------------------------------------------- J/K tile schedule
constexpr int JK_KN = 16;      // k-range per warp task
constexpr int JK_CHUNK = 64;   // l-chunk per warp task (2 per lane, one 128-bit load)
constexpr int JK_NI = 4;       // rows i per row-block
constexpr int JK_NJ = 2;       // rows j per row-block

struct RowBlock {
    int i0, j0;
};
struct JKTask {
    int k0;        // first k of the range (multiple of JK_KN)
    int chunk;     // l-chunk index (l0 = 64 * chunk)
    int rb_begin;  // first row-block (index into the row-block list)
    int rb_count;
};
struct JKSchedule {
    int ni = 0, nj = 0;
    std::vector<RowBlock> h_rowblocks;
    std::vector<JKTask> h_tasks;
    RowBlock *d_rowblocks = nullptr;
    JKTask *d_tasks = nullptr;
    int n_tasks = 0;
};

struct Timer {
    cudaEvent_t a = nullptr, b = nullptr;
    bool used = false;
};

}  // namespace sqcc

struct sqcc_ctx {
    int device = 0;
    int sm_count = 0;
    cudaStream_t stream = nullptr;
    int64_t launches = 0;

    // ERI shard (caller-owned)
    double *d_eri = nullptr;
    int n = 0;
    int ldd = 0;  // padded leading dimension of the D/J~/K~ work matrices (multiple of 4)
    int64_t ij_begin = 0, ij_end = 0;
    int i_begin = 0, i_end = 0;  // i range covered by the shard (rows (i, *))
    int64_t native_len = 0;
    int64_t n_unique = 0;
    int64_t *d_rowoff = nullptr;  // [npair_local + 1] offsets relative to shard start
    int64_t *d_ioff = nullptr;    // [n + 1] offsets of rows (i,0) in the full native tensor
    int64_t shard_base = 0;       // offset of the shard's first row in the full native tensor
    sqcc::JKSchedule sched[3];    // row-block shapes [0]: 4x2 (RHF, J-only)  [1]: 2x2 (UHF)  [2]: 4x4 (tuning)
    int jk_variant = 0;
    int jk_cfg = 0;  // tuning knob of the pipelined kernel (warps, stages), see launch_pipe

    // J/K scratch
    double *d_dwork = nullptr;   // [3][N][ldd]: D2 = 2*(Da+Db), Da, Db (padded)
    double *d_acc = nullptr;     // [3][N][ldd]: J~, K~a, K~b
    unsigned int *d_counter = nullptr;
    double *d_io = nullptr;      // device staging for *_host calls
    size_t io_bytes = 0;
    double *h_pin = nullptr;     // pinned staging
    size_t pin_bytes = 0;

    // grid
    const double *d_phi = nullptr;
    const double *d_w = nullptr;
    int64_t G = 0;
    int grid_n = 0;
    double *d_gridwork = nullptr;
    size_t gridwork_bytes = 0;

    // mp2 scratch
    double *d_mp2work = nullptr;
    size_t mp2work_bytes = 0;
    double *d_mp2out = nullptr;
    size_t mp2out_n = 0;

    sqcc::Timer timers[8];
};

namespace sqcc {
int ensure_io(sqcc_ctx *ctx, size_t bytes);
int ensure_pin(sqcc_ctx *ctx, size_t bytes);
int timer_start(sqcc_ctx *ctx, int slot);
int timer_stop(sqcc_ctx *ctx, int slot);

// eri_layout.cu
int eri_build_tables(sqcc_ctx *ctx);
int eri_synth(sqcc_ctx *ctx, uint64_t seed);
int eri_pack_full_slab(sqcc_ctx *ctx, const double *d_slab, int p0, int np);
int eri_pack_canonical(sqcc_ctx *ctx, const double *d_rows, int64_t ij0, int64_t nrows);
int eri_unpack_full(sqcc_ctx *ctx, double *d_full);

// jk_fused.cu
int jk_build_schedule(sqcc_ctx *ctx);
int jk_build(sqcc_ctx *ctx, const double *d_D, int nspin, int want_k, double *d_J, double *d_K);

// grid_lda.cu
int lda_vxc(sqcc_ctx *ctx, const double *d_D, int nspin, double *d_Vxc, double *d_Exc, double *d_rho);
int grid_quadrature(sqcc_ctx *ctx, const double *d_v, double *d_V);
int grid_rho(sqcc_ctx *ctx, const double *d_D, double *d_rho);

// mp2.cu
int ao2mo(sqcc_ctx *ctx, const double *d_C1, int n1, const double *d_C2, int n2, const double *d_C3, int n3,
          const double *d_C4, int n4, double *d_out);
int mp2(sqcc_ctx *ctx, const double *d_C, const double *d_eps, int n_occ, double *d_ecorr, double *d_iajb);
}  // namespace sqcc
