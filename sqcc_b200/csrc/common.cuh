// Shared declarations for libsqcc_b200.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>
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
    unsigned int mask;   // bit a*NJ+b: row (i0+a, j0+b) is held by this shard
    int pad;
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

// Peer group of the fused J/K exchange (comm.cu)
constexpr int COMM_MAXW = 16;
struct CommRank;
struct Comm {
    int world = 1, rank = 0;
    bool attached = false;      // peer_base[0..world) are valid
    bool enabled = true;        // sqcc_comm_set_enabled: false -> sqcc_jk_build returns partials even with a group attached
    bool ipc = false;           // peers were opened with cudaIpcOpenMemHandle (other processes)
    bool same_device = false;   // in-process group whose contexts share one device (single-launch emulation)
    double *d_sym = nullptr;    // symmetric region [acc 3 nl][res 3 nl][flags 2 COMM_MAXW u64]
    void *peer_base[COMM_MAXW] = {};
    unsigned long long epoch = 0;
    unsigned int bar_count = 0;
    unsigned int *d_gridbar = nullptr;
    unsigned long long *d_stamps = nullptr;   // [8] phase timestamps of the last exchange launch
    CommRank *d_view = nullptr;  // [COMM_MAXW]
    bool view_valid = false;     // d_view[0] holds this rank's view for outputs (view_J, view_K)
    double *view_J = nullptr, *view_K = nullptr;
    int *h_status = nullptr, *d_status = nullptr;   // mapped host word
};

struct ScfState;   // scf.cu

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

    // ERI shard (caller-owned).  A shard holds, for every i, the rows (i, j) with j in [jlo[i], jhi[i]) -- stored in
    // pair-index order.  Row-range shards [ij_begin, ij_end) and tile-range shards (sqcc_shard_tiles) are both
    // expressed this way.
    double *d_eri = nullptr;
    int n = 0;
    int ldd = 0;  // padded leading dimension of the D/J~/K~ work matrices (multiple of 4)
    std::vector<int> jlo, jhi;   // [n] held j interval per i (empty: jlo == jhi)
    std::vector<int64_t> h_rowptr;  // [n + 1] number of held rows with i' < i
    std::vector<int64_t> h_ioff;    // [n] local offset such that row (i,j) starts at h_ioff[i] + j Te(i) + Te(j)
    int64_t nrows = 0;           // held rows
    bool whole = false;          // every row of the tensor is held (unsharded)
    int i_begin = 0, i_end = 0;  // i range with held rows
    int64_t native_len = 0;
    int64_t n_unique = 0;
    int64_t *d_rowoff = nullptr;  // [nrows + 1] offsets of the held rows relative to the shard start
    int64_t *d_rowij = nullptr;   // [nrows] global pair index of the held rows
    int64_t *d_ioff = nullptr;    // [n] device copy of h_ioff
    int *d_jlo = nullptr, *d_jhi = nullptr;   // [n] device copies of jlo / jhi (one allocation)
    sqcc::JKSchedule sched[3];    // row-block shapes [0]: 4x2 (RHF, J-only)  [1]: 2x2 (UHF)  [2]: 4x4 (tuning)
    int jk_variant = 0;
    int jk_cfg = 0;  // tuning knob of the pipelined kernel (warps, stages), see launch_pipe

    // J/K scratch
    double *d_dwork = nullptr;   // [3][N][ldd]: D2 = 2*(Da+Db), Da, Db (padded)
    double *d_acc = nullptr;     // [3][N][ldd]: J~, K~a, K~b (first part of comm.d_sym)
    sqcc::Comm comm;
    unsigned int *d_counter = nullptr;
    unsigned long long *d_trace = nullptr;  // tuning: per-warp task trace of the last pipelined J/K launch (sqcc_jk_trace)
    int trace_warps = 0;
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

    sqcc::ScfState *scf = nullptr;   // device-resident SCF iteration state (scf.cu)

    sqcc::Timer timers[8];
};

namespace sqcc {
int ensure_io(sqcc_ctx *ctx, size_t bytes);
int ensure_pin(sqcc_ctx *ctx, size_t bytes);
int timer_start(sqcc_ctx *ctx, int slot);
int timer_stop(sqcc_ctx *ctx, int slot);

// eri_layout.cu
void pair_unrank_host(int64_t ij, int &i, int &j);
int64_t tile_count(int n);
int tile_intervals(int n, int64_t t_begin, int64_t t_end, std::vector<int> &jlo, std::vector<int> &jhi);
int row_intervals(int n, int64_t ij_begin, int64_t ij_end, std::vector<int> &jlo, std::vector<int> &jhi);
int64_t intervals_native_len(int n, const std::vector<int> &jlo, const std::vector<int> &jhi);
int eri_build_tables(sqcc_ctx *ctx);
int eri_synth(sqcc_ctx *ctx, uint64_t seed);
int eri_pack_full_slab(sqcc_ctx *ctx, const double *d_slab, int p0, int np);
int eri_pack_canonical(sqcc_ctx *ctx, const double *d_rows, int64_t ij0, int64_t nrows);
int eri_unpack_full(sqcc_ctx *ctx, double *d_full);
int eri_pack_blocks(sqcc_ctx *ctx, const double *d_values, const int64_t *d_desc, int nblk);

// jk_fused.cu
int jk_build_schedule(sqcc_ctx *ctx);
int jk_partial(sqcc_ctx *ctx, const double *d_D, int nspin, int want_k);
int jk_build(sqcc_ctx *ctx, const double *d_D, int nspin, int want_k, double *d_J, double *d_K);

// comm.cu
int comm_alloc(sqcc_ctx *ctx);
void comm_free(sqcc_ctx *ctx);
int jk_reduce(sqcc_ctx *ctx, int nspin, int want_k, double *d_J, double *d_K);

// grid_lda.cu
int lda_vxc(sqcc_ctx *ctx, const double *d_D, int nspin, double *d_Vxc, double *d_Exc, double *d_rho);
int grid_quadrature(sqcc_ctx *ctx, const double *d_v, double *d_V);
int grid_rho(sqcc_ctx *ctx, const double *d_D, double *d_rho);

// scf.cu
void scf_free(sqcc_ctx *ctx);

// mp2.cu
int ao2mo(sqcc_ctx *ctx, const double *d_C1, int n1, const double *d_C2, int n2, const double *d_C3, int n3,
          const double *d_C4, int n4, double *d_out);
int mp2(sqcc_ctx *ctx, const double *d_C, const double *d_eps, int n_occ, double *d_ecorr, double *d_iajb);
}  // namespace sqcc
