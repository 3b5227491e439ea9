/*
 * openrdm_b200.h -- C ABI of the B200-native MC-PDFT energy engine.
 *
 * This is the drop-in boundary for OpenRDM's MC-PDFT energy hot path
 *     mcpdft::mcpdft_energy()                   /root/reference/src/libmcpdft/energy.cc:27-188
 *     mcpdft::MCPDFT::build_rho/build_pi/build_R/translate   src/libmcpdft/mcpdft.cc:35-416
 *     mcpdft::Functional::EX_LSDA/EC_VWN3/EX_PBE/EC_PBE       src/libfunc/functional.cc:12-309
 * OpenRDM itself has no C ABI on this path (it is a C++ static library, CMakeLists.txt:23-24);
 * the host-side C++ mirror of its classes (openrdm_b200/host/) and any other binding sit on
 * top of exactly these entry points.  Plain pointers and sizes only; no CUDA / torch types.
 *
 * Conventions (the reference's): all matrices are column-major FP64;
 *   phi(p,mu) at phi[p + mu*ld]            (arma::mat(npts,nbfs), mcpdft.h:184)
 *   D1(mu,nu) at D1[mu + nu*n]
 *   D2ab is (n*n) x (n*n), element (nu*n+mu, sigma*n+lambda) multiplies
 *        phi_mu phi_nu phi_lambda phi_sigma  (mcpdft.cc:204; build_tpdm = kron(D1a,D1b), :640-652)
 *
 * Every function returns ORDM_OK (0) or an error code; ordm_last_error() gives the text.
 * There is no CPU fallback anywhere: without a CUDA device every compute call fails loudly.
 * A context is bound to one device and is not thread-safe; use one context per host thread.
 */
#ifndef OPENRDM_B200_H
#define OPENRDM_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct ordm_ctx ordm_ctx;

enum ordm_status {
  ORDM_OK = 0,
  ORDM_ERR_INVALID = 1, /* bad argument */
  ORDM_ERR_STATE = 2,   /* call order: inputs missing for the requested stage */
  ORDM_ERR_CUDA = 3,    /* CUDA runtime error (no device, launch failure, out of memory) */
  ORDM_ERR_NCCL = 4,    /* NCCL not loadable / communicator error */
  ORDM_ERR_UNSUPPORTED = 5
};

/* energy.cc:114-120 dispatches on the string "SVWN"; every other string means PBE.  REVPBE, BLYP
 * and BOP are the Psi4 plugin's further choices (MCPDFT_FUNCTIONAL, src/libpsi4/mcpdft_solver.cc:742-787):
 * EX_PBE_I(kappa = 1.245) + EC_PBE_I, EX_B88_I + EC_LYP_I, EX_B88_I + EC_B88_OP. */
enum ordm_functional { ORDM_FUNC_SVWN = 0, ORDM_FUNC_PBE = 1, ORDM_FUNC_REVPBE = 2, ORDM_FUNC_BLYP = 3, ORDM_FUNC_BOP = 4,
                       ORDM_FUNC_WPBE = 5 /* EX_wPBE_I + EC_PBE_I (:747-751), omega from ordm_set_omega */ };

/* 0-3: the four libfunc entry points (functional.h:17-43).  4-7: the plugin's routines
 * (src/libpsi4/lsda_gga_functionals.cc: EX_PBE_I with kappa = 1.245 :317-398, EX_B88_I :207-273,
 * EC_LYP_I :764-829, EC_B88_OP :478-518 -- the last has no density threshold in the reference and is
 * NaN wherever a spin density is zero; so it is here). */
enum ordm_xc_term { ORDM_EX_LSDA = 0, ORDM_EC_VWN3 = 1, ORDM_EX_PBE = 2, ORDM_EC_PBE = 3,
                    ORDM_EX_REVPBE = 4, ORDM_EX_B88 = 5, ORDM_EC_LYP = 6, ORDM_EC_B88_OP = 7,
                    ORDM_EX_WPBE = 8, /* EX_wPBE_I, src/libpsi4/rs_functionals.cc:86-300 (short-range omega-PBE exchange) */
                    ORDM_EC_PW92 = 9  /* EC_PW92_I, lsda_gga_functionals.cc:997-1093 */ };

/* Per-point vectors held by the engine = the arma::vec members of mcpdft::MCPDFT
 * (mcpdft.h:216-305) that the path reads or writes. */
enum ordm_vector {
  ORDM_VEC_RHOA = 0, ORDM_VEC_RHOB, ORDM_VEC_RHO,
  ORDM_VEC_RHOA_X, ORDM_VEC_RHOA_Y, ORDM_VEC_RHOA_Z,
  ORDM_VEC_RHOB_X, ORDM_VEC_RHOB_Y, ORDM_VEC_RHOB_Z,
  ORDM_VEC_SIGMA_AA, ORDM_VEC_SIGMA_AB, ORDM_VEC_SIGMA_BB,
  ORDM_VEC_PI, ORDM_VEC_R, ORDM_VEC_TR_RHOA, ORDM_VEC_TR_RHOB,
  ORDM_VEC_TR_SIGMA_AA, ORDM_VEC_TR_SIGMA_AB, ORDM_VEC_TR_SIGMA_BB,
  ORDM_VEC_PI_X, ORDM_VEC_PI_Y, ORDM_VEC_PI_Z,
  ORDM_VEC_W,
  ORDM_VEC_COUNT
};

/* option bits for ordm_set_options */
enum {
  ORDM_OPT_DIAGNOSTICS = 1, /* keep every per-point vector (getters work); off = energy-only mode */
  ORDM_OPT_PI_GRADIENT = 2, /* also build grad Pi (MCPDFT::build_ontop_pair_density_gradients,
                               mcpdft.cc:214-270; needs phi_x/y/z).  mcpdft_energy runs that stage for
                               every GGA case although translate() never reads its result; here it is
                               on request, and all three components are stored (the reference drops
                               pi_y, mcpdft.cc:268) */
  ORDM_OPT_FULLY_TRANSLATED = 4, /* ordm_energy / ordm_energy_host use MCPDFT::fully_translate
                               (mcpdft.cc:418-606) instead of MCPDFT::translate; with orbital gradients
                               this implies ORDM_OPT_PI_GRADIENT (the ft gradients consume grad Pi) */
  ORDM_OPT_SERIAL_STAGES = 8 /* ordm_energy runs density -> on-top -> XC back to back on one stream
                               (per-kernel timing); default: the HBM-bound density kernel streams
                               underneath the tensor-bound on-top kernel */
};

/* What mcpdft_energy returns or prints: energy.cc:175-183, mcpdft.cc:100-104,339-343. */
typedef struct {
  double e_tot, ex, ec, eclass;
  double n_tot, n_a, n_b;       /* integrated total / alpha / beta density */
  double trn_tot, trn_a, trn_b; /* integrated translated densities */
  /* CUDA-event times of the last ordm_energy call on this context's stream, milliseconds */
  double ms_density, ms_pi, ms_xc, ms_reduce, ms_total;
  uint64_t npts;          /* points this context (shard) processed */
  uint32_t gpu_launches;  /* kernels launched by that call */
  uint32_t overlapped;    /* 1: the density kernel ran underneath the on-top kernel (ms_density and
                             ms_pi overlap in time); 2: only its core-orbital sums did, the active-space
                             part followed (ms_density = both parts); 0: stages ran back to back */
  uint32_t pi_guard_points; /* > 0: the int8 tensor-core on-top kernel's error guard flagged that many points
                             (its a-posteriori estimate exceeded 1e-12 max(1, |Pi|)), and the whole call was
                             repeated with the FP64 tensor-pipe kernels -- which then stay selected until the
                             RDMs or the orbitals change; this result is the FP64 one */
  uint32_t density_tiled; /* launches of the TMA-tiled density kernel (k1_tile.cuh) in this call; 0: the thread-per-point
                             k1_density served it (diagnostic vectors requested, no orbital gradients, > 28 active orbitals ...) */
} ordm_result;

/* Which kernel computes the active-space on-top pair density for the RDMs currently set (no reference
 * counterpart; for benchmarks and logs).  *int8_slice_pairs > 0: the int8 tensor-core kernel
 * (tcgen05.mma.kind::i8 over exact byte slices of 46-bit fixed-point images, csrc/k2_int8.cuh; 18..20 active
 * orbitals, Pi without grad Pi) with that many slice-pair products; 0: the FP64 tensor-pipe kernels
 * (DMMA.8x8x4, csrc/k2_ontop.cuh).  Environment ORDM_K2_INT8 = 0 | 1 (26 pairs, default) | 2 (21 pairs) picks
 * at context creation.  Error model of the int8 path (k2_int8.cuh, guard_flag): the 46-bit images and the dropped
 * low-order slice pairs leave, per point, an error of standard deviation
 *     sqrt( 2^(4 em) g1 S2^2 + g2 S2^4 ) / 4,   S2 = sum_t phi_t^2,  2^em > max_t |phi_t|,
 * g1, g2 from the 2-RDM (ordm::int8_guard_constants); the kernel evaluates 4 standard deviations for every point
 * against the parity bar 1e-12 max(1, |Pi|).  If any point exceeds it the call is repeated with the FP64 kernels
 * (ordm_result.pi_guard_points), so the int8 path never returns a Pi it cannot vouch for; with 21 pairs (mode 2, opt-in)
 * the dropped-pair term is ~300x larger and the guard trips correspondingly earlier.  *mma_per_tile = tensor-core instructions per 128-point (int8) / 32-point (FP64) tile.
 * Pointers may be NULL; ORDM_ERR_STATE before the 2-RDM is set. */
int ordm_pi_kernel_info(const ordm_ctx* ctx, int* int8_slice_pairs, uint64_t* mma_per_tile);

/* ---- lifecycle ----------------------------------------------------------------------- */
const char* ordm_version(void);
int ordm_device_count(void);
/* replaces `new MCPDFT(...)` / `delete mc` (tests/main.cc:17,46) */
int ordm_create(int device, ordm_ctx** out);
void ordm_destroy(ordm_ctx* ctx);
/* ctx may be NULL: returns the last error of the calling thread's most recent failing call */
const char* ordm_last_error(const ordm_ctx* ctx);
/* use an existing cudaStream_t (passed as void*) instead of the context's own stream */
int ordm_set_stream(ordm_ctx* ctx, void* cuda_stream);
int ordm_set_options(ordm_ctx* ctx, unsigned flags);
/* MCPDFT_OMEGA of the Psi4 plugin (src/libpsi4/plugin.cc:50, default 0.0): the range-separation parameter
 * read by EX_wPBE_I (ORDM_FUNC_WPBE, ORDM_EX_WPBE) */
int ordm_set_omega(ordm_ctx* ctx, double omega);
int ordm_synchronize(ordm_ctx* ctx);

/* page-locked host memory, so that the H2D copies of set_* / energy_host run at PCIe speed */
int ordm_host_alloc(size_t bytes, void** out);
int ordm_host_free(void* p);

/* ---- inputs (host pointers; copied, never retained) ---------------------------------- */
/* MCPDFT::set_npts + set_w (mcpdft.h:112,114); reader: read_grids_from_file, utility.cc:5-34 */
int ordm_set_grid(ordm_ctx* ctx, size_t npts, const double* w);
/* MCPDFT::set_nbfs + set_phi/_x/_y/_z (mcpdft.h:113,118-121).  phi_x/y/z NULL = LDA only. */
int ordm_set_orbitals(ordm_ctx* ctx, size_t npts, int n, size_t ld, const double* phi,
                      const double* phi_x, const double* phi_y, const double* phi_z);
/* AO -> MO transform of the orbital values on the device, the step
 * directly in front of the path: TransformPhiMatrixAOMO (src/libpsi4/mcpdft_solver.cc:619-635, two
 * F_DGEMMs per irrep) / mo_values = np.matmul(ao_values, C_mo) (src/libpyscf/libpyscf.py:608).
 * ao, ao_x, ao_y, ao_z: npts x nao column-major host arrays (leading dimension ld; gradients NULL =
 * LDA); C: nao x nmo column-major MO coefficients.  The AO values are streamed through the device in
 * point batches (H2D overlapped with the FP64 GEMM of the previous batch: this library's own DMMA kernel,
 * csrc/k0_ao2mo.cuh -- no cuBLAS) and phi = ao . C (and its gradients) become the context's resident orbitals,
 * exactly as if ordm_set_orbitals had been called with the transformed matrices. */
int ordm_set_orbitals_ao(ordm_ctx* ctx, size_t npts, int nao, size_t ld, const double* ao, const double* ao_x,
                         const double* ao_y, const double* ao_z, int nmo, const double* C);
/* MCPDFT::set_D1a + set_D1b (mcpdft.h:125-126).  n x n, not assumed symmetric (mcpdft.cc:148). */
int ordm_set_rdm1(ordm_ctx* ctx, int n, const double* D1a, const double* D1b);
/* the D2ab argument of mcpdft_energy / build_pi (energy.h:14-18, mcpdft.h:35); dense
 * (n^2)x(n^2) in the reference's index convention, n = all orbitals of ordm_set_orbitals. */
int ordm_set_rdm2ab_dense(ordm_ctx* ctx, int n, const double* D2ab);
/* Extension (no reference counterpart): the orbitals are
 * [ncore doubly-occupied | nact active]; A1a/A1b are the nact x nact active 1-RDM blocks and
 * D2act the (nact^2)x(nact^2) active alpha-beta 2-RDM.  Equivalent to the dense calls above
 * with D1 = I (+) A and the core-expanded D2ab, without ever forming the (n^2)^2 matrix. */
int ordm_set_active_space(ordm_ctx* ctx, int ncore, int nact, const double* A1a, const double* A1b,
                          const double* D2act);

/* Sparse ingestion of the active-space RDMs -- what production callers
 * hand over instead of dense matrices:
 *   upper_packed = 0  Psi4 plugin lists  struct opdm {i,j,val} / struct tpdm {i,j,k,l,val}
 *                     (src/libpsi4/mcpdft_solver.h:71-83): each element contributes val * phi_i phi_j
 *                     (* phi_k phi_l) once, as BuildRhoFast / BuildPiFast do (mcpdft_solver.cc:1867-2061,
 *                     1578-1735);
 *   upper_packed = 1  PySCF exporter COO groups /SP_SYMM_D1/ACT_D1{a,b}_MO/{VAL,ROW_IDX,COL_IDX} and
 *                     /SP_SYMM_D2/ACT_D2ab_MO/{VAL,DIM1_IDX..DIM4_IDX} (src/libpyscf/libpyscf.py:246-463):
 *                     only i <= j (and k <= l) of the symmetric arrays is stored, values unscaled.
 * Indices are active-space local (0..nact-1); ncore doubly-occupied orbitals precede them, as in
 * ordm_set_active_space.  Cost O(nnz): the dense (nact^2)^2 matrix is never formed. */
int ordm_set_active_space_sparse(ordm_ctx* ctx, int ncore, int nact,
                                 size_t nnz1a, const int* i1a, const int* j1a, const double* v1a,
                                 size_t nnz1b, const int* i1b, const int* j1b, const double* v1b,
                                 size_t nnz2, const int* i2, const int* j2, const int* k2, const int* l2,
                                 const double* v2, int upper_packed);
/* Host-only helper: the same 2-RDM list expanded to the dense (n^2)x(n^2) reference convention
 * (mcpdft.cc:204) that ordm_set_rdm2ab_dense / mcpdft_energy take.  D2ab_out is overwritten. */
int ordm_rdm2ab_sparse_to_dense(int n, size_t nnz, const int* i2, const int* j2, const int* k2, const int* l2,
                                const double* v2, int upper_packed, double* D2ab_out);

/* MCPDFT::set_rhoa ... set_tr_sigma_bb, set_pi, set_pi_x/_y/_z, set_R (mcpdft.h:125-153): install a caller-provided
 * per-point vector (npts host doubles, copied).  What the reference's later stages read from its members is read
 * from the installed vector here too: rho and Pi by build_R / translate (mcpdft.cc:276-283,297-330), R by translate
 * and fully_translate (:321, :460) until the next ordm_build_R recomputes it, the spin-density gradient components
 * by translate_density_gradients (:383-398; needs ORDM_OPT_DIAGNOSTICS so that both spins of a direction exist),
 * grad Pi by fully_translate (:541-590).  The remaining vectors (rho_a, rho_b, sigma, translated quantities) only
 * feed the getters, as in the reference.  ordm_energy / ordm_energy_host rebuild every vector from the inputs, as
 * mcpdft_energy does.  Call ordm_set_options before installing vectors (changing the options re-allocates them). */
int ordm_set_vector(ordm_ctx* ctx, int which, const double* host_in, size_t npts);

/* ---- stages = the public methods of mcpdft::MCPDFT ------------------------------------ */
int ordm_build_rho(ordm_ctx* ctx);   /* MCPDFT::build_rho,  mcpdft.cc:35-40  */
int ordm_build_pi(ordm_ctx* ctx);    /* MCPDFT::build_pi,   mcpdft.cc:178-183 (D2ab from set_rdm2ab_*) */
int ordm_build_R(ordm_ctx* ctx);     /* MCPDFT::build_R,    mcpdft.cc:272-285 */
int ordm_translate(ordm_ctx* ctx);   /* MCPDFT::translate,  mcpdft.cc:287-292 */
/* MCPDFT::fully_translate, mcpdft.cc:418-423 (after build_rho, build_pi with ORDM_OPT_PI_GRADIENT
 * when the orbitals carry gradients, build_R) */
int ordm_fully_translate(ordm_ctx* ctx);
/* the integrated densities the reference prints after build_rho / translate */
int ordm_integrated_densities(ordm_ctx* ctx, double out_n[3], double out_trn[3]);

/* Functional::EX_LSDA / EC_VWN3 / EX_PBE / EC_PBE on caller-provided vectors
 * (functional.h:17-43): npts-long host arrays, weights from ordm_set_grid.  Unused sigma
 * arguments may be NULL. */
int ordm_functional(ordm_ctx* ctx, int term, size_t npts, const double* rho_a, const double* rho_b,
                    const double* sigma_aa, const double* sigma_ab, const double* sigma_bb,
                    double* energy_out);

/* ---- input files (north_star "Data loading (libio)") -------------------------------
 * The grid + orbitals + RDM container of the reference's ecosystem is the PySCF exporter's data.h5
 * (src/libpyscf/libpyscf.py:613-690: groups /N, /C, /E, /D/D1, /D/D2, /GRIDS/{W,X,Y,Z}, /PHI/PHI_{AO,MO}{,_X,_Y,_Z},
 * /SP_SYMM_D1, /SP_SYMM_D2 with the upper-packed COO RDMs :246-463); libio's own opdm.h5 / tpdm.h5 (src/libio/HDF5_Read_Compact.cc:8-12)
 * are a round trip whose result mcpdft_energy discards (energy.cc:132-140).  ordm_file_open reads that schema from
 * HDF5 itself when this library was built with <hdf5.h> available, and from the raw little-endian container with the
 * same dataset names (format: csrc/rawio.h; writer: openrdm_b200/rawio.py) otherwise.  Arrays are row-major (numpy). */
typedef struct ordm_file ordm_file;
enum { ORDM_DT_F64 = 0, ORDM_DT_I64 = 1, ORDM_DT_I32 = 2 };
int ordm_file_open(const char* path, ordm_file** out);
void ordm_file_close(ordm_file* f);
/* dataset by its data.h5 path, e.g. "/GRIDS/W": element type, rank, dims (padded with 1), pointer valid until close */
int ordm_file_dataset(ordm_file* f, const char* name, int* dtype, int* ndim, uint64_t dims[4], const void** data);
/* Everything the path needs from one file into the context: weights /GRIDS/W; orbitals on the grid from /PHI/PHI_MO
 * (first N_COR + N_CAS_ORB columns; transposed to the engine's column-major layout on the device) or, with
 * ORDM_LOAD_PREFER_AO or when PHI_MO is absent, from /PHI/PHI_AO and /C through the device AO -> MO transform;
 * gradients when all of _X, _Y, _Z are there; the active-space RDMs from the upper-packed COO groups
 * /SP_SYMM_D1/ACT_D1{a,b}_MO, /SP_SYMM_D2/ACT_D2ab_MO (or, with ORDM_LOAD_DENSE_RDM / when those are absent, from
 * /D/D1/ACT_D1{A,B}_MO and /D/D2/ACT_D2AB_MO); N_COR doubly occupied orbitals precede the N_CAS_ORB active ones.
 * Afterwards ordm_energy etc. run exactly as after the ordm_set_* calls.  info (nullable) reports what was found;
 * e_core + e_hartree (+ V_nn, which the exporter only prints) is the classical energy `eclass` of mcpdft_energy. */
enum { ORDM_LOAD_PREFER_AO = 1, ORDM_LOAD_DENSE_RDM = 2, ORDM_LOAD_NO_GRADIENTS = 4 };
typedef struct {
  uint64_t npts;
  int nao, nmo, ncore, nact;
  int gga, used_ao, used_sparse_rdm;
  double e_core, e_hartree;     /* /E/E_CORE, /E/E_HARTREE (libpyscf.py:629-630) */
} ordm_data_info;
int ordm_load_data(ordm_ctx* ctx, const char* path, unsigned flags, ordm_data_info* info);

/* ---- the Psi4 plugin's extras around the path -------------------------- */
/* MCPDFTSolver::polyradical_analysis (src/libpsi4/mcpdft_solver.cc:3001-3240): densities of effectively unpaired
 * electrons  U(r), D(r), N(r) = sum_ij phi_i phi_j {U, D, N}_ij  with, element by element on n = D1a + D1b,
 * U = 1 - |1 - n|, D = n (2 - n), N = n^2 (2 - n)^2 (:3104-3108), on the resident grid, orbitals and 1-RDMs
 * (doubly occupied core orbitals contribute nothing).  traces = sum_p w_p {U, D, N}(r_p) (:3185-3187); U_r, D_r, N_r:
 * optional npts-long host arrays (what the plugin writes to mygrids.txt, :3213-3227).  At most 32 active orbitals. */
int ordm_polyradical(ordm_ctx* ctx, double traces[3], double* U_r, double* D_r, double* N_r, size_t npts);

/* Energy assembly of MCPDFTSolver::compute_energy (mcpdft_solver.cc:849-925): the scalars the plugin takes from Psi4
 * (reference energy, V_nn, kinetic and nuclear-attraction energies of the 1-RDMs, classical Coulomb energy, and for
 * the double hybrids the HF-exchange and MP2-correlation energies) together with Ex, Ec of ordm_energy:
 *   two_electron = e_reference - V_nn - (T + V_ne);   wf = (T + V_ne) + lambda two_electron;
 *   dft = (1 - lambda)(J + Ex) + (1 - lambda^2) Ec;   E = wf + dft + V_nn  (+ lambda E_x^HF + lambda^2 E_c^MP2 for 1DH_MCPDFT).
 * lambda = MCPDFT_LAMBDA (plugin.cc:51); it is ignored (0) for method MCPDFT, as in the plugin (:717-737).  The
 * range-separated methods use the same assembly with the functional evaluated at MCPDFT_OMEGA (ordm_set_omega +
 * ORDM_FUNC_WPBE).  Host arithmetic only; no context needed. */
enum ordm_method { ORDM_METHOD_MCPDFT = 0, ORDM_METHOD_1H_MCPDFT, ORDM_METHOD_1DH_MCPDFT, ORDM_METHOD_RS_MCPDFT,
                   ORDM_METHOD_RS1H_MCPDFT, ORDM_METHOD_RS1DH_MCPDFT };   /* MCPDFT_METHOD, plugin.cc:47 */
typedef struct {
  double e_reference;            /* reference (CASSCF / v2RDM) total energy */
  double e_nuclear_repulsion, e_kinetic, e_nuclear_attraction;
  double e_coulomb;              /* classical Coulomb energy J */
  double e_hf_exchange, e_mp2_correlation;   /* 1DH_MCPDFT only */
} ordm_hybrid_inputs;
typedef struct { double e_total, wf_contribution, dft_contribution, one_electron, two_electron, lambda; } ordm_hybrid_result;
int ordm_hybrid_energy(int method, double lambda, const ordm_hybrid_inputs* in, double ex, double ec, ordm_hybrid_result* out);

/* ---- the path in one call -------------------------------------------------------------- */
/* mcpdft::mcpdft_energy (energy.cc:27-188) on the device-resident inputs: density -> Pi ->
 * R/translate/XC -> quadrature (-> NCCL all-reduce when a communicator is attached). */
int ordm_energy(ordm_ctx* ctx, int functional, double eclass, ordm_result* out);

/* Stream-ordered form of ordm_energy: ordm_energy_async enqueues the whole step (kernels; the last one does the final
 * passes, the all-reduce and stores the result scalars into the context's pinned result block) on the context's stream
 * and returns; ordm_energy_wait blocks
 * for the LAST enqueued step and returns its result.  Several steps may be enqueued back to back -- an SCF driver
 * that pipelines, a benchmark loop -- so that the host runs ahead of the GPU and the launch latency of a step
 * disappears from the device timeline (at 8 GPUs a step is < 1 ms).  If the int8 on-top kernel's guard flagged
 * points in any enqueued step, ordm_energy_wait repeats the last one with the FP64 kernels before returning.
 * Inputs must not be changed between an async call and its wait; ordm_energy refuses to run while a step is pending. */
int ordm_energy_async(ordm_ctx* ctx, int functional, double eclass);
int ordm_energy_wait(ordm_ctx* ctx, ordm_result* out);

/* Same, from HOST buffers: the grid is streamed through the device in point batches with
 * the H2D copies overlapped with compute; nothing but the RDMs has to be resident.  This is
 * the end-to-end call a host-memory caller (arma::mat) makes.  ncore/nact/A1a/A1b/D2act as
 * in ordm_set_active_space (ncore = 0 and nact = n for the plain dense case). */
int ordm_energy_host(ordm_ctx* ctx, int functional, double eclass, size_t npts, int n, size_t ld,
                     const double* w, const double* phi, const double* phi_x, const double* phi_y,
                     const double* phi_z, size_t batch_points, ordm_result* out);

/* ---- outputs --------------------------------------------------------------------------- */
/* MCPDFT::get_rhoa ... get_tr_sigma_bb (mcpdft.h:64-110): copies npts doubles to the host.
 * Needs ORDM_OPT_DIAGNOSTICS, except RHO / PI / W which always exist after their stage. */
int ordm_get_vector(ordm_ctx* ctx, int which, double* host_out, size_t npts);
/* MCPDFT::get_phi / get_phi_x / _y / _z (mcpdft.h:72-75): component 0..3, column-major into
 * host_out with leading dimension ld_host >= npts. */
int ordm_get_orbitals(ordm_ctx* ctx, int component, double* host_out, size_t ld_host);
/* raw device pointer of a per-point vector (NULL if not materialised) */
void* ordm_device_vector(ordm_ctx* ctx, int which);

/* ---- synthetic inputs of BASELINE.json's configs 3-5 (ordm_synth.h) ------------------- */
/* Fill this context's phi (and grad phi if gga) and w on the device for global points
 * [p0, p0+npts_local) of a grid of npts_total points. */
int ordm_synth_fill_device(ordm_ctx* ctx, uint64_t seed, size_t npts_total, size_t p0,
                           size_t npts_local, int n, int gga);
/* Same values into host arrays (column-major, ld = npts_local); any pointer may be NULL. */
int ordm_synth_fill_host(uint64_t seed, size_t npts_total, size_t p0, size_t npts_local, int n,
                         double* w, double* phi, double* phi_x, double* phi_y, double* phi_z);
/* N-representable ensemble-of-determinants active-space RDMs:
 * A1a, A1b nact x nact; D2act (nact^2)x(nact^2); nelec_a/b active electrons per spin. */
int ordm_synth_rdms(uint64_t seed, int nact, int nelec_a, int nelec_b, int ndet, double* A1a,
                    double* A1b, double* D2act);

/* ---- multi-GPU: one process per GPU, grid sharded by point ranges ---------------------- */
/* NCCL is loaded lazily (dlopen libnccl.so.2).  Rank 0 makes the id, the launcher ships the
 * 128 bytes to the other ranks (torch.distributed / MPI / a file), every rank attaches. */
int ordm_comm_unique_id(char id_out[128]);
int ordm_comm_attach(ordm_ctx* ctx, int nranks, int rank, const char id[128]);
/* geometry of the attached communicator; *peer_allreduce = 1 when the eight result scalars are
 * combined by this library's own NVLink peer-store kernel (exchange buffers of all ranks mapped through
 * CUDA IPC at attach time, rank-ordered and therefore bitwise identical sums on every rank), 0 when
 * ncclAllReduce is used (mapping not possible, or ORDM_NO_PEER_ALLREDUCE set).  Pointers may be NULL. */
int ordm_comm_info(const ordm_ctx* ctx, int* nranks, int* rank, int* peer_allreduce);
int ordm_comm_detach(ordm_ctx* ctx);
/* contiguous shard [p0, p0+count) of rank `rank`: balanced in units of 64 points */
void ordm_shard_range(size_t npts_total, int nranks, int rank, size_t* p0, size_t* count);

#ifdef __cplusplus
}
#endif
#endif /* OPENRDM_B200_H */
