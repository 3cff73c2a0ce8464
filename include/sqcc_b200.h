/* sqcc_b200.h -- C ABI of libsqcc_b200.so: the B200-native SCF Fock-build hot path of SQCC.
 *
 * The reference (QuantumChemistrySchoolJapan-HFT/sqcc) has NO plugin / FFI seam for this path: the
 * contractions are Python loop nests inlined in hf_ksdft.Calculator.scf() and mp2.Calculator.mp2().
 * Each entry point below therefore cites the reference loop nest it replaces (file:line under
 * /root/reference/code/python3).  INTEGRATION.md shows the ctypes binding a maintainer of the reference
 * would add at those lines.
 *
 * Conventions
 *   - every function returns 0 on success, a negative sqcc_status on error; sqcc_last_error() returns
 *     a thread-local message.  No exceptions cross the ABI.
 *   - one sqcc_ctx drives ONE device (one process per GPU).  A context is not thread-safe.  The cross-rank sum of the
 *     partial J/K of a sharded tensor is done inside sqcc_jk_build once a peer group is attached (sqcc_comm_*); without
 *     one the caller all-reduces the output buffers itself (torch.distributed / NCCL).
 *   - "d_" pointers are device pointers owned by the CALLER (e.g. torch tensors' data_ptr()); the
 *     library never frees them.  "h_" pointers are host pointers.  Scratch lives in the context.
 *   - all matrices are row-major float64.  D / J / K / Vxc are dense [N,N] (or [nspin,N,N]).
 *   - work is enqueued on the context's stream (sqcc_ctx_set_stream; default: the legacy stream 0).
 *     Functions with "_host" in the name synchronise that stream before returning.
 *
 * Packed ERI layouts (8-fold symmetry, i>=j, k>=l, ij>=kl, ij = i(i+1)/2+j):
 *   canonical : offset ij(ij+1)/2 + kl, plain (ij|kl).
 *   native v2 : tile-interleaved (sqcc_b200/csrc/layout.cuh; restated in oracle/packing.py).  A tile ("pair-block")
 *               is a 4 x 4 block of rows (i, j); inside a tile, for every sub-row k and 64-column chunk c, the 16 rows
 *               lie side by side as one contiguous "piece" [16][w] (w = 64, or 16/32/48 on the triangular edge
 *               l <= k), pieces ordered chunk-major; cells outside the canonical range are stored as 0.0.  The stored
 *               value is (ij|kl) * 0.5^[i==j] * 0.5^[k==l] * 0.5^[ij==kl].  A rank holds a contiguous tile range
 *               [t_begin, t_end) ("pair-block shard") = a contiguous byte range of the tensor.
 */
#ifndef SQCC_B200_H
#define SQCC_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct sqcc_ctx sqcc_ctx;

typedef enum {
    SQCC_OK = 0,
    SQCC_ERR_INVALID = -1,   /* bad argument / state */
    SQCC_ERR_CUDA = -2,      /* CUDA runtime error (message in sqcc_last_error) */
    SQCC_ERR_NOMEM = -3,
    SQCC_ERR_UNSUPPORTED = -4
} sqcc_status;

/* ---- library / context --------------------------------------------------------------------- */
int sqcc_version(void);
const char *sqcc_last_error(void);
int sqcc_device_count(int *count);
int sqcc_ctx_create(int device, sqcc_ctx **out);
int sqcc_ctx_destroy(sqcc_ctx *ctx);
int sqcc_ctx_set_stream(sqcc_ctx *ctx, void *cuda_stream);
int sqcc_ctx_synchronize(sqcc_ctx *ctx);

/* ---- layout helpers (pure host arithmetic, usable without a GPU) --------------------------- */
int64_t sqcc_npair(int n);
int64_t sqcc_nquad(int n);
/* Tile-granular shards.  A tile ("pair-block") is a 4 x 4 block of rows (i, j): i-blocks are anchored at the top of
 * the i range (the last block ends at i = N), j-blocks at multiples of 4; global tile order is i-block ascending,
 * j-block ascending.  A shard is a contiguous tile range [t_begin, t_end), cut for equal streaming
 * cost; every 4 x 4 row-block of the J/K kernel is whole on exactly one rank. */
int64_t sqcc_tile_count(int n);
int sqcc_shard_tiles(int n, int world, int rank, int64_t *t_begin, int64_t *t_end);
/* Same with unequal shares: rank r gets weights[r] / sum(weights) of the streaming cost (relative speeds of the GPUs:
 * under their power caps the GPUs of one box sustain different clocks; equal shards leave the fast ones waiting). */
int sqcc_shard_tiles_weighted(int n, int world, int rank, const double *weights, int64_t *t_begin, int64_t *t_end);
/* Rows held by the tile range: for every i the interval [jlo[i], jhi[i]) of j (arrays of N ints). */
int sqcc_tile_rows(int n, int64_t t_begin, int64_t t_end, int *jlo, int *jhi);
/* Length in doubles of the native-v2 shard holding tiles [t_begin, t_end). */
int64_t sqcc_native_len_tiles(int n, int64_t t_begin, int64_t t_end);
/* Offset (doubles, relative to the start of tile t_begin) of element (i,j,k,l) in native v2; -1 if not canonical or
 * stored before t_begin. */
int64_t sqcc_native_offset(int n, int64_t t_begin, int i, int j, int k, int l);

/* ---- ERI ingest: replaces the consumption of interface_psi4.py:177-190 (full N^4 host tensor) ---- */
/* Attach a caller-owned device buffer of sqcc_native_len_tiles(...) doubles (16-byte aligned) as this rank's ERI shard
 * and build the task schedule.  The buffer must stay alive until detach/destroy.  N <= 4096 (64-bit offsets
 * throughout; what fits is bounded by HBM: N = 494 is 63.5 GB, N ~ 640 fills one 180 GB B200, N ~ 1080 eight). */
int sqcc_eri_attach_tiles(sqcc_ctx *ctx, double *d_native, int n, int64_t t_begin, int64_t t_end);
/* Fill the attached shard with the counter-based synthetic tensor (oracle/synth.py). */
int sqcc_eri_synth(sqcc_ctx *ctx, uint64_t seed);
/* Pack a slab ERI[p0:p0+np, :, :, :] of the full tensor (device pointer, [np,N,N,N]) into the shard. */
int sqcc_eri_pack_full_slab(sqcc_ctx *ctx, const double *d_slab, int p0, int np);
/* Same, from a pageable/pinned HOST tensor ERI[N,N,N,N]; streams slabs through a pinned staging buffer. */
int sqcc_eri_pack_full_host(sqcc_ctx *ctx, const double *h_eri_full);
/* Pack canonical packed rows [ij0, ij0+nrows) (device pointer to the first element of row ij0). */
int sqcc_eri_pack_canonical(sqcc_ctx *ctx, const double *d_rows, int64_t ij0, int64_t nrows);
/* Direct ingestion without the N^4 host tensor: dense blocks (i0..i0+ni) x (j0..) x (k0..) x
 * (l0..), row-major [ni][nj][nk][nl] -- e.g. one per canonical shell quartet, Psi4: MintsHelper.ao_eri_shell(M,N,P,Q).
 * d_desc holds nine int64 per block: offset into d_values, i0, ni, j0, nj, k0, nk, l0, nl.  Every element goes to its
 * canonical position if this shard holds its row; call sqcc_eri_clear first (pad cells must be zero). */
int sqcc_eri_clear(sqcc_ctx *ctx);
int sqcc_eri_pack_blocks(sqcc_ctx *ctx, const double *d_values, const int64_t *d_desc, int nblk);
/* Expand the attached (unsharded) ERI into the full [N,N,N,N] device tensor (plain values). */
int sqcc_eri_unpack_full(sqcc_ctx *ctx, double *d_eri_full);

/* ---- fused Coulomb/exchange build: replaces hf_ksdft.py:483-495 (RHF/RKS) and :510-535 (UHF/UKS) ----
 * nspin = 1: J = J(D), K = K(D);  nspin = 2: J = J(D[0]+D[1]), K[s] = K(D[s]).
 * want_k = 0 skips exchange (KS-DFT, hf_ksdft.py:488/:517).  One pass over this rank's ERI shard; the
 * outputs are this rank's PARTIAL (already symmetrised) J/K -- sum over ranks gives the full matrices -- unless a
 * peer group is attached (sqcc_comm_attach below), in which case they are the full matrices on every rank. */
int sqcc_jk_build(sqcc_ctx *ctx, const double *d_D, int nspin, int want_k, double *d_J, double *d_K);
/* Same through host buffers (H2D of D, D2H of J/K inside; stream-synchronous).  Page-locked host buffers are copied
 * from / to directly; pageable ones go through the context's pinned staging area. */
int sqcc_jk_build_host(sqcc_ctx *ctx, const double *h_D, int nspin, int want_k, double *h_J, double *h_K);
/* Select the kernel: 0 = streaming warp-task kernel, TMA ring + narrow-piece pass (default), 1 = simple atomics kernel
 * (cross-check), 16+c = streaming kernel, tuning configuration c kept for A/B measurements (1: cp.async ring instead of
 * TMA, 2: no L2 cache-policy hints, 6: spin-split UHF pass, 7: one pass over all tasks; other values = default). */
int sqcc_jk_set_variant(sqcc_ctx *ctx, int variant);
/* Deterministic mode (on != 0): every flush of a partial sum into the J~/K~ work matrices is converted to 64-bit fixed
 * point and added with an INTEGER atomic (associative), so J and K are bit-for-bit reproducible from run to run, under
 * any task schedule, and -- with the fused exchange, which sums the ranks in a fixed order -- identical on all ranks.
 * The scale is 2^s with (max |stored integral|) * max(sum |D2|, sum |Da|, sum |Db|) * 2^s < 2^59: resolution
 * ~1e-15 relative to that bound (parity gate: 1e-11 absolute).  Ranks of a group must share one integral bound:
 * sqcc_eri_absmax(ctx, &v, 0) reads this shard's, sqcc_eri_absmax(ctx, &v, 1) sets the agreed one (max over ranks). */
int sqcc_jk_set_deterministic(sqcc_ctx *ctx, int on);
int sqcc_eri_absmax(sqcc_ctx *ctx, double *absmax, int set);

/* ---- cross-rank exchange of the partial J/K over NVLink peer memory --------------------------
 * With a peer group attached, sqcc_jk_build / sqcc_jk_build_host return the FULL J/K on every rank: one kernel per
 * rank sums the partial accumulators of all ranks through mapped peer memory (reduce-scatter + all-gather in one pass,
 * fixed summation order -> bit-identical results on all ranks) and symmetrises them into the caller's buffers.
 * Without a group the outputs are this rank's partials and the caller sums them (e.g. torch.distributed all_reduce).
 *   multi-process (one process per GPU): after sqcc_eri_attach*, every rank calls sqcc_comm_export, the 64-byte
 *     handles are all-gathered by the host (torch.distributed), then every rank calls sqcc_comm_attach.
 *   one process, several contexts: sqcc_comm_attach_local(ctxs, world); contexts on distinct devices then use
 *     sqcc_jk_build as above; contexts sharing ONE device (tests) must run sqcc_jk_partial on each and finish with a
 *     single sqcc_jk_reduce_local launch (kernels that wait for each other cannot be separate launches on one GPU).
 * Every rank of a group must issue the same sequence of builds.  sqcc_comm_status: < 0 after a barrier timeout
 * (the timed-out rank publishes nothing, tells its peers through the barrier flags, and every rank's J/K of that build
 * are NaN), else the group size (0: none). */
#define SQCC_COMM_HANDLE_BYTES 64
int sqcc_comm_export(sqcc_ctx *ctx, void *handle /* SQCC_COMM_HANDLE_BYTES */);
int sqcc_comm_attach(sqcc_ctx *ctx, int world, int rank, const void *handles /* world * SQCC_COMM_HANDLE_BYTES */);
int sqcc_comm_attach_local(sqcc_ctx **ctxs, int world);
int sqcc_comm_detach(sqcc_ctx *ctx);
int sqcc_comm_status(sqcc_ctx *ctx);
/* The status word is sticky: after a failed exchange (every rank's J/K are NaN-filled) clear it before reuse. */
int sqcc_comm_reset_status(sqcc_ctx *ctx);
/* enabled = 0: sqcc_jk_build returns this rank's partials although a group is attached (A/B against an all-reduce). */
int sqcc_comm_set_enabled(sqcc_ctx *ctx, int enabled);
/* Tuning aid: globaltimer (ns) at the phase boundaries of the last exchange launch on this rank: start, after barrier A,
 * after the reduce+broadcast, grid barrier done, after barrier B, end (out8[6..7] unused). */
int sqcc_comm_stamps(sqcc_ctx *ctx, uint64_t *out8);
/* prep + fused kernel only: leaves the partial accumulators in the context (no J/K output). */
int sqcc_jk_partial(sqcc_ctx *ctx, const double *d_D, int nspin, int want_k);
int sqcc_jk_reduce_local(sqcc_ctx **ctxs, int world, int nspin, int want_k, double **d_J, double **d_K);

/* ---- LDA exchange grid quadrature: replaces hf_ksdft.py:426-440, :634-646 (rho), ksdft_functional.py
 * :41-65, :68-95, :98-134, :137-177 (v_x, E_x) and hf_ksdft.py:456-461, :469-476 (V_xc) ---- */
int sqcc_grid_attach(sqcc_ctx *ctx, const double *d_phi /*[G,N]*/, const double *d_w /*[G]*/, int64_t G, int n);
/* Vxc [nspin,N,N], Exc [1], rho [nspin,G] (optional, may be NULL); all device pointers. */
int sqcc_lda_vxc(sqcc_ctx *ctx, const double *d_D, int nspin, double *d_Vxc, double *d_Exc, double *d_rho);
int sqcc_lda_vxc_host(sqcc_ctx *ctx, const double *h_D, int nspin, double *h_Vxc, double *h_Exc, double *h_rho);
/* Generic quadrature V_pq = sum_n phi_np (w_n v_n) phi_nq for a caller-supplied v (QM/MM external
 * potential, interface_psi4.py:399-403) and rho on the attached points for a given D. */
int sqcc_grid_quadrature(sqcc_ctx *ctx, const double *d_v /*[G]*/, double *d_V /*[N,N]*/);
int sqcc_grid_rho(sqcc_ctx *ctx, const double *d_D /*[N,N]*/, double *d_rho /*[G]*/);

/* ---- MP2: replaces mp2.py:89-99 (denominators), :123-170 (quarter transforms), :240-254 (energy) ----
 * C is [N,N] AO x MO (columns = MOs, first n_occ occupied), eps [N].  Needs an unsharded ERI attached.
 * d_iajb (optional) receives (ia|jb) as [n_occ,n_virt,n_occ,n_virt]. */
int sqcc_mp2(sqcc_ctx *ctx, const double *d_C, const double *d_eps, int n_occ, double *d_ecorr, double *d_iajb);
int sqcc_mp2_host(sqcc_ctx *ctx, const double *h_C, const double *h_eps, int n_occ, double *h_ecorr,
                  double *h_iajb);
/* Generic AO->MO transform (pq|rs) -> (ab|cd) with four coefficient blocks [N,n1..n4] (CIS needs (ij|ab),
 * cis_tdhf_tda_tddft.py:128-211).  Output [n1,n2,n3,n4]. */
int sqcc_ao2mo(sqcc_ctx *ctx, const double *d_C1, int n1, const double *d_C2, int n2, const double *d_C3, int n3,
               const double *d_C4, int n4, double *d_out);

/* ---- device-resident SCF iteration ------------------------------------------------------
 * Replaces, between two Fock builds, the host steps hf_ksdft.py:498-504 / :527-548 (Fock), :552-587 (electronic
 * energy), :97-123 + :604-621 (X^T F X, eigh, C = X C') and :150-156 / :178-199 + :625 (density): D, J/K, V_xc, F, C
 * stay in HBM, one double (E_elec) is read back per iteration.  eigh is cuSOLVER Dsyevd (dlopen'ed on first use).
 *   setup        H_core [N,N], X = S^-1/2 [N,N] (host), nspin = 1 restricted / 2 unrestricted, occupations, ksdft flag
 *   set_density  initial density [nspin,N,N] (host)
 *   fock_energy  J/K (+ LDA V_xc) of the resident density, Fock matrices, *h_e_elec = electronic energy
 *   update       new orbitals and density from the current Fock matrices (the reference skips this after convergence)
 *   get          density [nspin,N,N], C^T [nspin,N,N] (row k = MO k), orbital energies [nspin,N] to the host */
int sqcc_scf_setup(sqcc_ctx *ctx, const double *h_H, const double *h_X, int nspin, int n_occ_a, int n_occ_b, int ksdft);
int sqcc_scf_set_density(sqcc_ctx *ctx, const double *h_D);
int sqcc_scf_fock_energy(sqcc_ctx *ctx, double *h_e_elec);
/* The two halves of sqcc_scf_fock_energy for a sharded tensor whose partial J/K are summed by the CALLER (NCCL
 * all-reduce instead of the fused peer exchange): sqcc_scf_jk leaves [J; K...] of the resident density in a
 * library-owned device buffer (*d_jk, *n_doubles doubles; this rank's partials when nothing sums them), the caller
 * all-reduces it in place, sqcc_scf_finish builds the Fock matrices and the energy from it.  sqcc_scf_fock_energy
 * itself refuses a sharded tensor without an active peer exchange (it would use partial matrices). */
int sqcc_scf_jk(sqcc_ctx *ctx, void **d_jk, int64_t *n_doubles);
int sqcc_scf_finish(sqcc_ctx *ctx, double *h_e_elec);
int sqcc_scf_update(sqcc_ctx *ctx);
int sqcc_scf_get(sqcc_ctx *ctx, double *h_D, double *h_Ct, double *h_eps);

/* ---- instrumentation ------------------------------------------------------------------------ */
/* CUDA-event duration (ms) of the dominant kernel(s), MEAN over the calls since the previous query (at most the last
 * 64, so a timed back-to-back loop reports its own in-loop kernel time): slot 0 = J/K kernel, 1 = J/K finalize /
 * exchange, 2 = rho, 3 = vxc, 4 = mp2 transforms, 5 = mp2 energy, 6 = SCF update (X^T F X, eigh, density); -1 for
 * slots without a call since the last query.  Returns the number of slots written. */
int sqcc_get_timings(sqcc_ctx *ctx, double *ms, int n);
/* Number of library kernels launched since context creation. */
int64_t sqcc_launch_count(sqcc_ctx *ctx);
/* Bytes of the attached native shard (actual) and of its unique integrals (algorithmic, 8 per integral). */
int sqcc_eri_bytes(sqcc_ctx *ctx, int64_t *native_bytes, int64_t *algorithmic_bytes);

#ifdef __cplusplus
}
#endif
#endif /* SQCC_B200_H */
