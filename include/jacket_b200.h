/* jacket_b200.h -- C ABI of libjacket_b200.so (B200 / sm_100a only).
 *
 * The reference (JK-hqy/Small-FEM-Solver-based-on-a-lot-of-assumption,
 * JacketAnalysisGUI_v2.py, cited as GUI.py:LINE) has no FFI: its "operator
 * interface" for the hot path is the Python class surface
 *   MorisonCalculator.compute_all_morison_forces / find_critical_phase  (GUI.py:591-724)
 *   FEMSolver.__init__ / apply_boundary_conditions / solve /
 *             get_reactions / get_member_internal_forces                (GUI.py:438-533)
 * called from JacketAnalysisGUI.run_analysis (GUI.py:1827-2082).  The Python
 * package in this repo keeps that surface and forwards to the entry points
 * below through ctypes (INTEGRATION.md shows the binding).
 *
 * Conventions
 *   - plain pointers and sizes only; every array is caller-owned HOST memory
 *     unless the name ends in _dev;
 *   - every function returns 0 on success, a negative JK_E* code on failure;
 *     jk_last_error(h) (or jk_last_error(NULL) for jk_create failures) gives
 *     the message;
 *   - one opaque handle per (structure, GPU); a handle owns one CUDA stream
 *     (or borrows the one passed to jk_create) and all its device buffers;
 *   - calls are blocking unless stated; there is NO CPU fallback: without a
 *     CUDA device jk_create fails with JK_ENODEVICE.
 *   - units follow the reference: coordinates m, section properties mm,
 *     forces N, moments N*mm, displacements mm / rad, stresses MPa.
 */
#ifndef JACKET_B200_H
#define JACKET_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct jk_handle_s* jk_handle_t;

#define JK_OK            0
#define JK_EINVAL       -1
#define JK_ECUDA        -2
#define JK_ENODEVICE    -3
#define JK_ENOTSPD      -4   /* Cholesky met a non-positive pivot (the reference would fall to lstsq, GUI.py:486-487) */
#define JK_ESTATE       -5   /* call order violated (e.g. scan before factor) */

/* per-section property row passed to jk_create (mm units, GUI.py:122-137) */
#define JK_SEC_NPROP     8
#define JK_SEC_D_OUTER   0
#define JK_SEC_AX        1
#define JK_SEC_IY        2
#define JK_SEC_IZ        3
#define JK_SEC_IX        4
#define JK_SEC_AY        5
#define JK_SEC_AZ        6
#define JK_SEC_R_OUTER   7

/* columns of the per-phase result table written by jk_phase_scan / jk_morison_scan.
 * 0..7 are the reference's find_critical_phase row (GUI.py:705-714); column 1
 * (phase_deg) = numpy.degrees(omega*t) % 360 evaluated on the device with the same
 * IEEE operations (two multiplies, exact fmod), bit-identical to the reference.
 * 8..15 are the per-phase FEM summary (run_analysis replayed at t_i). */
#define JK_TABLE_NCOL       16
#define JK_COL_T             0
#define JK_COL_PHASE_DEG     1
#define JK_COL_TOTAL_KN      2
#define JK_COL_DRAG_KN       3
#define JK_COL_INERTIA_KN    4
#define JK_COL_FX_KN         5
#define JK_COL_FY_KN         6
#define JK_COL_FZ_KN         7
#define JK_COL_MAX_DISP_MM   8   /* max nodal |translation|, GUI.py:2035-2040 */
#define JK_COL_MAX_DISP_NODE 9   /* node index of that maximum (first maximum) */
#define JK_COL_MAX_UTIL     10   /* max member utilisation, GUI.py:2054 */
#define JK_COL_MAX_UTIL_MEM 11   /* member index of that maximum (first maximum) */
#define JK_COL_MAX_VM_MPA   12
#define JK_COL_SUM_RX       13   /* sum of support reactions, N (GUI.py:2027-2033) */
#define JK_COL_SUM_RY       14
#define JK_COL_SUM_RZ       15

/* member result row (GUI.py:521-532 numeric fields, in this order) */
#define JK_MEMBER_NCOL       7   /* Fx_max_kN Fy_max_kN Fz_max_kN My_max_kNm Mz_max_kNm von_mises_max_MPa utilization */
/* Morison member detail row (GUI.py:668-674) */
#define JK_DETAIL_NCOL       4   /* drag_kN inertia_kN total_kN submerged_length */

/* node ordering of the free-free system */
#define JK_ORDER_NATURAL     0   /* reference order (node_list order) */
#define JK_ORDER_RCM         1   /* reverse Cuthill-McKee on the member graph (minimises the tile band) */
/* solver storage */
#define JK_SOLVER_BANDED     0   /* tile band derived from the ordering */
#define JK_SOLVER_DENSE      1   /* full lower triangle (the reference's dense K_ff, GUI.py:482) */

/* stage timers returned by jk_get_timings (milliseconds, CUDA events on the handle's stream) */
#define JK_NTIMERS          13
#define JK_T_ASSEMBLE        0
#define JK_T_FACTOR          1
#define JK_T_WAVE_SETUP      2
#define JK_T_MORISON         3
#define JK_T_RHS             4
#define JK_T_SOLVE_FWD       5
#define JK_T_SOLVE_BWD       6
#define JK_T_POST            7
#define JK_T_REDUCE          8
#define JK_T_SCAN_TOTAL      9
#define JK_T_H2D            10
#define JK_T_D2H            11
#define JK_T_SOLVE_FWD2     12   /* forward sweep launches issued after the factor join (split factor); -1 when not split */

int         jk_version(void);
const char* jk_last_error(jk_handle_t h);

/* Documented run-time options of a handle (the library never reads the environment).  Defaults in brackets.
 * Scheduling switches -- every setting produces bit-identical results (tests/test_gpu_parity.py):
 *   start_gate [1]          the main stream waits until the factor clusters are resident (asynchronous factorisation)
 *   start_gate2 [1]         the first forward sweep parts wait until the second factor segment is resident
 *   post_overlap [1]        member post of first-chain chunks beside the second chain's backward sweep
 *   early_totals [1]        Morison columns of the table reduced on a side stream behind the Morison kernel
 *   cuda_graph [1]          jk_step / jk_step_dev replay the whole step (assemble + factor + scan) as a captured CUDA graph
 *   fused_loads [0]         1: the Morison kernel lumps member end forces into nodal loads itself (register runs + deposit
 *                           rows, finalised in-kernel by a chained look-back; halves the load stage's HBM traffic but
 *                           measured slower at c4, 1.81 vs 1.70 ms); 0: member forces are written to HBM and gathered by a
 *                           second kernel in the reference's member order.  Same sums in another fixed order (1e-15)
 *   gather_blocks [1]       phase blocks of the Morison stage: the load gather of block b (HBM-bound, side stream) runs beside
 *                           the Morison kernel of block b + 1 (FP64-bound); gather_rows [0] = grid rows of those gather launches
 *   sweep_slab [0]          right-hand sides per triangular-sweep CTA: 0 = chosen so that the CTAs fill the SMs, 8, 16, 32
 * Ordering / storage switches, read by the next jk_set_supports (results agree to rounding, the reference's run_analysis
 * has no counterpart: GUI.py:481-490 is a dense LU):
 *   two_chains [1]  factor_split [1]  split_pct [70]  level_regroup [1]  support_rooted_rcm [1]  tma_sweep [1]
 *   blocked_inverse [1]
 * Debug aids: profile_chol [0], profile_sweep [0] (1: clock breakdown of the sweep warps on stderr, 2: plus per-item clock
 * stamps), debug_factor_delay [0] (clocks), debug_morison_smem_pad [0] (KB, occupancy probe), debug_fuse_mode [0] (timing
 * probes of the fused load path; 1 and 2 give WRONG results by design).
 * jk_option_count / jk_option_name enumerate the keys. */
int         jk_set_option(jk_handle_t h, const char* key, int value);
int         jk_get_option(jk_handle_t h, const char* key, int* value);
int         jk_option_count(void);
const char* jk_option_name(int index);

/* Replaces CustomJacketStructure (GUI.py:302-354) as the geometry carrier:
 * xyz[n_nodes*3] (m, z = 0 at MWL), conn[n_members*2] node indices,
 * sec_id[n_members] rows of sec_props[n_sec*JK_SEC_NPROP].
 * stream: a cudaStream_t to borrow, or NULL to create one. */
int jk_create(int device, void* stream,
              int n_nodes, const double* xyz,
              int n_members, const int32_t* conn, const int32_t* sec_id,
              int n_sec, const double* sec_props,
              jk_handle_t* out);
int jk_destroy(jk_handle_t h);

/* FEMSolver.apply_boundary_conditions (GUI.py:473-479): all 6 DOF of each
 * fixed node are removed.  Chooses the ordering / storage of K_ff. */
int jk_set_supports(jk_handle_t h, int n_fixed, const int32_t* fixed_nodes, int ordering, int solver);

/* FEMSolver.__init__ (GUI.py:439-467): element stiffness (BeamElement3D,
 * GUI.py:361-422), transform, deterministic assembly of K_ff into tile storage. */
int jk_assemble(jk_handle_t h, double E, double G);
/* factor half of np.linalg.solve (GUI.py:485): blocked Cholesky, once per structure */
int jk_factor(jk_handle_t h);
/* Asynchronous variant: queued on the handle's side stream; the next scan overlaps its Morison + load stage with the
 * factorisation and joins before the triangular sweeps.  A non-positive pivot is reported by that scan's
 * jk_phase_scan / jk_read_table / jk_solve (JK_ENOTSPD). */
int jk_factor_begin(jk_handle_t h);

/* F_global contributions that do not depend on the phase (interface loads
 * GUI.py:1962-1977 and self-weight GUI.py:1994-2012), built by the host. */
int jk_set_static_load(jk_handle_t h, const double* F_static /* [6*n_nodes] */);

/* RaschiiWave closed-form branch (GUI.py:265, 277-281) */
int jk_set_wave_airy(jk_handle_t h, double a, double k, double omega, double d, double U_c, double dt);
/* Fourier-series kinematics (Stokes / Fenton form, wrapper semantics GUI.py:259-281):
 * eta = sum E[j] cos(j phi) ; u = sum B[j] cosh(j k zb)/cosh(j k d) cos(j phi) + U_c ; j = 1..n_harm */
int jk_set_wave_fourier(jk_handle_t h, double k, double omega, double d, double U_c, double dt,
                        int n_harm, const double* E, const double* B);
/* MorisonCalculator.__init__ (GUI.py:544-557) + the Gauss rule of GUI.py:615-617
 * (nodes s in [0,1] and weights w, computed by the host with numpy.leggauss). */
int jk_set_morison(jk_handle_t h, double theta_wave, double theta_current,
                   double rho, double Cd, double Cm,
                   int n_gauss, const double* gauss_s, const double* gauss_w);

/* MorisonCalculator.find_critical_phase (GUI.py:684-724): Morison only.
 * table[P*JK_TABLE_NCOL] (columns 8.. are 0), *critical = first index of max total_kN. */
int jk_morison_scan(jk_handle_t h, int P, const double* t, double* table, int64_t* critical);

/* MorisonCalculator.compute_all_morison_forces at one t (GUI.py:591-682):
 * nodal_forces[n_nodes*3], totals[9] = drag xyz, inertia xyz, morison xyz (N),
 * details[n_members*JK_DETAIL_NCOL]; any output may be NULL. */
int jk_morison_single(jk_handle_t h, double t, double* nodal_forces, double* totals, double* details);

/* MorisonCalculator.get_kinematics_3d (GUI.py:559-589, with RaschiiWave.get_kinematics GUI.py:290-296 inside) for n
 * points xyz[n*3] at time t, evaluated by the device functions of the Morison kernels:
 * out[n*10] = u_wave v_wave w_wave u_current v_current du_dt dv_dt dw_dt submerged(0/1) eta. */
int jk_kinematics_points(jk_handle_t h, int n, const double* xyz, double t, double* out);

/* The whole hot path for P phases: Morison -> RHS -> two triangular sweeps ->
 * reactions / member forces / utilisation -> per-phase table -> critical phase.
 * Full per-phase results stay in HBM; fetch rows with jk_fetch_phase.
 * With table == NULL and critical == NULL the call only queues the work (host times in, nothing read back, no
 * synchronisation): read the results later with jk_read_table or on the device through jk_table_dev. */
int jk_phase_scan(jk_handle_t h, int P, const double* t, double fy, double* table, int64_t* critical);
/* Same, but nothing is copied: t_dev[P] is already in HBM and the table stays
 * there (jk_read_table copies it out).  Asynchronous on the handle's stream. */
int jk_phase_scan_dev(jk_handle_t h, int P, const double* t_dev, double fy);
int jk_read_table(jk_handle_t h, int P, double* table, int64_t* critical);

/* One whole analysis step for a resident loop (design iterations, ensembles of structures): jk_assemble(E, G) +
 * jk_factor_begin + the phase scan, as one call.  With option cuda_graph = 1 (default) the step's ~30 launches on three
 * streams are captured once into a CUDA graph and replayed while nothing baked into it has changed (supports, wave, Morison
 * set-up, moduli, fy, P, options, buffer sizes); results are bit-identical to the separate calls.
 * jk_step_dev: t_dev[P] already in HBM, nothing is read back, no host synchronisation (results: jk_read_table, jk_table_dev,
 * jk_critical_*_dev).  jk_step: host times in, host table + critical index out (NULL, NULL: queue only). */
int jk_step_dev(jk_handle_t h, double E, double G, int P, const double* t_dev, double fy);
int jk_step(jk_handle_t h, double E, double G, int P, const double* t, double fy, double* table, int64_t* critical);

/* Sea-state ensemble (BASELINE configs[4]): n_states Airy sea states (amplitude a = H/2, wave number k from the
 * dispersion relation GUI.py:197-206, omega = 2 pi / T, math heading theta_wave) x n_phase phases each, evaluated as
 * ONE batch of n_states*n_phase load cases on the factor already computed.  Depth, current, dt come from
 * jk_set_wave_airy, current heading / coefficients / Gauss rule from jk_set_morison.  t[n_states*n_phase] are the
 * case times (state-major).  F_dir (nullable) = two load vectors [2][6*n_nodes] added as cos(theta_wave)*F_dir[0] +
 * sin(theta_wave)*F_dir[1]: the interface shear that run_analysis applies along the wave heading (GUI.py:1967-1971).
 * table[n_states*n_phase*JK_TABLE_NCOL]; critical_per_state[n_states] = first phase index of the maximum total_kN
 * inside each state (GUI.py:717 applied per sea state).  Rows of any case: jk_fetch_phase. */
int jk_ensemble_scan(jk_handle_t h, int n_states, int n_phase, const double* a, const double* k, const double* omega,
                     const double* theta_wave, const double* t, const double* F_dir, double fy, double* table,
                     int64_t* critical_per_state);

/* The same ensemble from the raw sea-state parameters: significant numbers as the GUI takes them (H [m], T [s], wave
 * direction in compass degrees, GUI.py:1880-1900).  Amplitude, omega, the dispersion Newton iteration (GUI.py:197-206, one
 * thread per sea state, the reference's start value and stopping test), the heading cos/sin (GUI.py:548) and the case times
 * t = i*T/n_phase (GUI.py:696) are computed on the device; nothing but 3*n_states doubles crosses the bus.
 * k_out[n_states] (nullable) receives the wave numbers. */
int jk_ensemble_scan_sea_states(jk_handle_t h, int n_states, int n_phase, const double* H, const double* T,
                                const double* wave_dir_deg, double gravity, const double* F_dir, double fy,
                                double* table, int64_t* critical_per_state, double* k_out);

/* FEMSolver.solve with caller-built right-hand sides (GUI.py:481-490):
 * F[nrhs*6*n_nodes] (row = load case) -> same pipeline as jk_phase_scan minus Morison. */
int jk_solve(jk_handle_t h, int nrhs, const double* F, double fy);

/* Full rows of one phase / load case of the last scan or solve:
 * U[6*n_nodes] (GUI.py:488-490), reactions[6*n_fixed] (GUI.py:492-502),
 * member_rows[n_members*JK_MEMBER_NCOL] (GUI.py:521-532),
 * end_forces[n_members*12] (GUI.py:427-432, node-1 sign already flipped),
 * nodal_forces[n_nodes*3] Morison nodal loads (NULL-able, zeros after jk_solve). */
int jk_fetch_phase(jk_handle_t h, int phase, double* U, double* reactions,
                   double* member_rows, double* end_forces, double* nodal_forces);
/* One member-result column for all phases of the last scan: out[P] */
int jk_fetch_member_column(jk_handle_t h, int member, int column, int P, double* out);

/* Introspection used by the FEMSolver facade and the parity tests */
int jk_get_dims(jk_handle_t h, int32_t* out /* [12]: n_nodes n_members n_fixed n_free_dof n_pad tile band_tiles n_tiles
                                                 dof_half_bandwidth n_chains separator_tile_row separator_nodes */);
int jk_get_order(jk_handle_t h, int32_t* free_nodes /* [n_nodes-n_fixed] solver order */);
int jk_get_K(jk_handle_t h, double* K /* [n_dof*n_dof] dense, reference DOF order */);
int jk_get_elements(jk_handle_t h, double* Ke /* [M*144] */, double* Kl /* [M*144] */, double* R /* [M*9] */, double* L /* [M] */);
int jk_get_timings(jk_handle_t h, double* ms /* [JK_NTIMERS] */);
/* max_i |K u - F|_i / max_i |F|_i over the free DOFs of every phase of the last scan (diagnostic) */
int jk_residual(jk_handle_t h, double* rel_residual);
/* Solver statistics after a factorisation: out[0] = non-zeros of L (exact count, lower triangle incl. diagonal),
 * out[1] = FP64 flops the two triangular sweeps EXECUTE per load case (DMMA k-groups kept by the zero-block masks,
 * or all tile products of the band on the legacy path), out[2] = sweep items per slab (0 on the legacy path),
 * out[3] = 1 if the TMA / mbarrier sweep pipeline is active, 0 for the cp.async slab sweep,
 * out[4] = non-zeros of L for the candidate ordering with the smallest envelope (symbolic count; the ordering in use
 * trades a larger envelope for fewer mask blocks), out[5] = right-hand sides per sweep CTA of the last solve,
 * out[6] = state of the step graph (1: a captured CUDA graph is being replayed by jk_step*, 0: none, -1: capture failed on
 * this driver and the steps are launched eagerly), out[7] = kernels inside that graph. */
int jk_solver_stats(jk_handle_t h, double* out /* [8] */);
/* Host-only introspection of the sweep item list (no device needed; used by the CPU tests): the program the TMA
 * sweep runs for a chain of n_tiles tile rows, tile half-bandwidth band_tiles, first partial / known tile row kx
 * (= n_tiles for a plain sweep), first_tile[n_tiles] (nullable) = first tile column of every tile row's envelope (tiles
 * left of it are not visited; the tile next to the diagonal always is).  items[6*i..] = row, src, flags, xinfo, next_row, next_init; meta[3] = first known
 * row, known rows preloaded, top row of the ring numbering.  Returns the item count (items may be NULL to size). */
int jk_sweep_program(int n_tiles, int band_tiles, int kx, int backward, const int32_t* first_tile, int32_t* items, int cap_items,
                     int32_t* meta);
/* kernels launched by this handle since creation (bench "gpu_launches") */
int64_t jk_launch_count(jk_handle_t h);
void* jk_stream(jk_handle_t h);
/* device pointers of the last scan's table [P*JK_TABLE_NCOL] and of the (max value, first index) pair
 * produced by the on-device argmax -- for device-to-device collectives (NCCL) without a host round trip */
void* jk_table_dev(jk_handle_t h);
void* jk_critical_value_dev(jk_handle_t h);
void* jk_critical_index_dev(jk_handle_t h);

#ifdef __cplusplus
}
#endif
#endif
