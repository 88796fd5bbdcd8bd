/*
 * mvs_b200.h — C ABI of the B200-native non-rigid alignment hot path of MultiViewStitch.
 *
 * The reference (zjuzly/MultiViewStitch) has no FFI/plugin layer: its hot path sits behind the
 * public C++ methods of SRTSolver / Alignment / Deformation / Model2Depth / PartRecognition that
 * only `Processor` calls.  This header is the C-ABI shared-library surface a
 * maintainer binds instead of those classes; the headers under `multiviewstitch_b200/shim/` wrap it back into
 * classes with the reference's names.  Every entry point cites the reference interface it
 * replaces as  [ref: path:line]  (paths relative to /root/reference/MultiViewStitch/).
 *
 * Conventions
 *  - every function returns an int status: MVS_OK (0) or a negative MVS_ERR_*; nothing ever calls
 *    exit() (the reference's error convention is `cerr << msg; exit(-1)`, e.g.
 *    Deformation/Deformation.cpp:22-26 — replaced by status codes + mvs_last_error()).
 *  - all matrices are row-major; points are packed xyz triples unless the name says `xyzw`
 *    (16-byte float4 records, w ignored on input).
 *  - `*_host` pointers are caller-owned host memory, `*_dev` pointers are caller-owned device
 *    memory on the context's device.  The callee copies on set_input / create exactly like the
 *    reference copies in SetInput (Solver/SRTSolver.h:12-20, Deformation/Deformation.cpp:34).
 *  - there is NO CPU fallback: without a CUDA device every compute call returns
 *    MVS_ERR_NO_DEVICE.
 */
#ifndef MVS_B200_H
#define MVS_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MVS_OK 0
#define MVS_ERR_INVALID_ARG (-1)
#define MVS_ERR_CUDA (-2)
#define MVS_ERR_NO_DEVICE (-3)
#define MVS_ERR_STATE (-4)      /* call order violated (e.g. estimate before set_input) */
#define MVS_ERR_CAPACITY (-5)   /* an internal fixed-capacity table overflowed */
#define MVS_ERR_DEGENERATE (-6) /* input degenerate for the requested operation (n == 0 ...) */
#define MVS_ERR_COMM (-7)       /* NCCL / communicator failure */

typedef struct mvs_ctx mvs_ctx;
typedef struct mvs_nn_index mvs_nn_index;
typedef struct mvs_srt mvs_srt;
typedef struct mvs_deform mvs_deform;

/* Pinhole camera, x_cam = R x_world + t, u = (int)(fx x/z + cx + 0.5).
 * [ref: Camera/Camera.h:44-49 (K,R,t,w,h), Camera/Camera.cpp:45-48,68-72] */
typedef struct mvs_camera {
  double K[9];
  double R[9];
  double t[3];
  int32_t w, h;
} mvs_camera;

/* ---------------------------------------------------------------- context ------------------ */
const char* mvs_last_error(void);
int mvs_version(void);
int mvs_device_count(int* count);
/* stream: a cudaStream_t to launch on (e.g. the caller's current stream) or NULL for a private
 * non-blocking stream. */
int mvs_ctx_create(mvs_ctx** ctx, int device, void* stream);
int mvs_ctx_destroy(mvs_ctx* ctx);
int mvs_ctx_sync(mvs_ctx* ctx);
int mvs_ctx_stream(mvs_ctx* ctx, void** stream);
/* Per-kernel-family device timing (CUDA events on the launching stream).  Off by default. */
int mvs_prof_enable(mvs_ctx* ctx, int on);
int mvs_prof_reset(mvs_ctx* ctx);
int mvs_prof_get(mvs_ctx* ctx, const char* name, double* total_ms, int64_t* launches);
/* '\n'-separated list of family names seen so far; returns needed size in *needed. */
int mvs_prof_names(mvs_ctx* ctx, char* buf, int64_t cap, int64_t* needed);
/* Number of kernels this library launched on ctx since creation / last reset. */
int mvs_launch_count(mvs_ctx* ctx, int64_t* launches);
int mvs_launch_count_reset(mvs_ctx* ctx);

/* ---------------------------------------------------------------- exact kNN ---------------- */
/* Exact k nearest neighbours in float32 squared-L2, ties broken by lowest index; results per
 * query sorted ascending by (d2, index).  Replaces every cv::flann::Index(KDTreeIndexParams)
 * + knnSearch call on the path:
 * [ref: PartRecognition/PartRecognition.cpp:53-74 (k=1), Deformation/Deformation.cpp:68-97 (k=16),
 *  :110-134 (k=9), :245-284 (k=1)].  d2 = (dx*dx + dy*dy) + dz*dz in round-to-nearest float32
 * with no fused multiply-add, so it is bit-reproducible on the host. k <= 32. */
int mvs_nn_index_create(mvs_ctx* ctx, mvs_nn_index** out);
int mvs_nn_index_destroy(mvs_nn_index* idx);
int mvs_nn_index_build_host(mvs_nn_index* idx, const float* pts_xyz, int64_t n);
int mvs_nn_index_build_host_f64(mvs_nn_index* idx, const double* pts_xyz, int64_t n);
int mvs_nn_index_build_dev(mvs_nn_index* idx, const float* pts_xyzw_dev, int64_t n);
int mvs_nn_knn_host(mvs_nn_index* idx, const float* q_xyz, int64_t nq, int k, int32_t* out_idx,
                    float* out_d2);
int mvs_nn_knn_host_f64(mvs_nn_index* idx, const double* q_xyz, int64_t nq, int k,
                        int32_t* out_idx, float* out_d2);
int mvs_nn_knn_dev(mvs_nn_index* idx, const float* q_xyzw_dev, int64_t nq, int k,
                   int32_t* out_idx_dev, float* out_d2_dev);
/* One-shot convenience (build + query + destroy), host buffers. */
int mvs_nn_search(mvs_ctx* ctx, const float* q_xyz, int64_t nq, const float* t_xyz, int64_t nt,
                  int k, int32_t* out_idx, float* out_d2);
/* 1-NN label transfer.  [ref: PartRecognition/PartRecognition.cpp:50-77 PartRecog] */
int mvs_part_recog(mvs_ctx* ctx, const double* template_xyz, int64_t n_template,
                   const int32_t* template_labels, const double* query_xyz, int64_t n_query,
                   int32_t* out_labels);

/* ---------------------------------------------------------------- SRT solver --------------- */
/* [ref: Solver/SRTSolver.h:8-39] */
int mvs_srt_create(mvs_ctx* ctx, mvs_srt** out);
int mvs_srt_destroy(mvs_srt* h);
/* [ref: Solver/SRTSolver.h:12-20 SetInput] p = matches[i].first, q = matches[i].second, n x 3
 * doubles each.  Kept in float64 on the device (matches come from double unprojection,
 * Processor.cpp:806-811). cameras may be NULL if residual/RANSAC are not used.  The callee COPIES:
 * p/q may be reused or freed as soon as the call returns, pinned or pageable (for a pinned source the
 * call waits for the transfer). */
int mvs_srt_set_input(mvs_srt* h, const double* p_xyz, const double* q_xyz, int64_t n,
                      const mvs_camera* cam1, const mvs_camera* cam2);
/* float32 xyzw (float4) variants: host or device resident (the C2 bandwidth sweep layout). */
int mvs_srt_set_input_f32(mvs_srt* h, const float* p_xyzw, const float* q_xyzw, int64_t n,
                          const mvs_camera* cam1, const mvs_camera* cam2);
int mvs_srt_set_input_dev_f32(mvs_srt* h, const float* p_xyzw_dev, const float* q_xyzw_dev,
                              int64_t n, const mvs_camera* cam1, const mvs_camera* cam2);
/* [ref: Solver/SRTSolver.h:21-22] */
int mvs_srt_set_iteration_num(mvs_srt* h, int iter_num);
int mvs_srt_set_print(mvs_srt* h, int on);
/* [ref: Solver/SRTSolver.cpp:272-275 EstimateTransform(double&,Matrix3d&,Vector3d&)]
 * closed form: EstimateScale (:31-46, mean of centroid-distance ratios) then EstimateRT (:65-129,
 * Kabsch with the |det R + 1| <= 1e-9 reflection fix). */
int mvs_srt_estimate(mvs_srt* h, double* scale, double R[9], double t[3]);
/* [ref: Solver/SRTSolver.cpp:277-280 EstimateTransformRansac / :131-185 EstimateRTRansac]
 * triples: iter_num x 3 match indices, one hypothesis each (the reference draws them with
 * Shuffle(idx,size,3) from unseeded rand(), Common/Utils.h:25-34 — not replayable, so the sample
 * is an input).  triples == NULL: drawn from a counter-based RNG with `seed`.
 * best_err / best_iter may be NULL. Strict-< argmin, first wins. */
int mvs_srt_estimate_ransac(mvs_srt* h, const int32_t* triples, uint64_t seed, double* scale,
                            double R[9], double t[3], double* best_err, int32_t* best_iter);
/* [ref: Solver/SRTSolver.cpp:256-270 EstimateTransform(double&,double[3][3],double[3])] = RANSAC
 * path with array outputs. */
int mvs_srt_estimate_array(mvs_srt* h, const int32_t* triples, uint64_t seed, double* scale,
                           double R[3][3], double t[3]);
/* [ref: Solver/SRTSolver.cpp:6-29 ResidualError] mean symmetric pixel reprojection error. */
int mvs_srt_residual(mvs_srt* h, double scale, const double R[9], const double t[3], double* err);
/* The triples the library would draw for (n, iter_num, seed) — lets a caller replay RANSAC. */
int mvs_srt_draw_triples(int64_t n, int iter_num, uint64_t seed, int32_t* triples);
/* Per-match symmetric reprojection test of the inlier loop [ref: Processor/Processor.cpp:210-239]:
 * flags[i] = 1 iff both pixel errors are <= pixel_thr under (scale, R, t); err_sum receives the sum
 * over ALL matches of (e1 + e2)/2 (the reference's err_all_).  flags: n bytes, host. */
int mvs_srt_inliers(mvs_srt* h, double scale, const double R[9], const double t[3], double pixel_thr,
                    uint8_t* flags, double* err_sum, int64_t* n_inliers);

/* [ref: Processor/Processor.cpp:177-269 RemoveOutliers] `rounds` (3) rounds of: 200-hypothesis
 * RANSAC on the surviving matches, symmetric reprojection test with the threshold
 * pixel_err * adapt_ratio^round, compaction.  keep_idx (capacity n) receives the indices of the
 * surviving matches in their original order; err = err_all / size of the last round;
 * inlier_ratio is 0 as in the reference (the code that would update it is commented out, :243-253).
 * Hypothesis samples come from the counter-based RNG with seed + round. */
typedef struct mvs_outlier_params {
  int32_t rounds;      /* 3   [ref: Processor.cpp:198] */
  int32_t iter_num;    /* 200 [ref: Processor.cpp:202] */
  double pixel_err;    /* ParamParser::pixel_err (55; config.txt 60) */
  double adapt_ratio;  /* ParamParser::adapt_pixel_err_ratio (0.6; config.txt 0.75) */
  uint64_t seed;
} mvs_outlier_params;
int mvs_outlier_default_params(mvs_outlier_params* p);
int mvs_remove_outliers(mvs_ctx* ctx, const double* p_xyz, const double* q_xyz, int64_t n,
                        const mvs_camera* cam1, const mvs_camera* cam2,
                        const mvs_outlier_params* params, int32_t* keep_idx, int64_t* n_keep,
                        double* err, double* inlier_ratio);
/* The same round loop with the hypothesis samples of every round supplied by the caller:
 * draw(size, iter_num, out_triples, user) is called once per round with the number of matches still
 * in play and fills iter_num x 3 indices < size.  The reference draws them from the C library's
 * rand() (Common/Utils.h:25-34 Shuffle, Solver/SRTSolver.cpp:152): a caller that wants the
 * reference's exact result hands that stream over here.  draw == NULL: as mvs_remove_outliers. */
typedef void (*mvs_triples_fn)(int64_t size, int iter_num, int32_t* out_triples, void* user);
int mvs_remove_outliers_samples(mvs_ctx* ctx, const double* p_xyz, const double* q_xyz, int64_t n,
                                const mvs_camera* cam1, const mvs_camera* cam2,
                                const mvs_outlier_params* params, mvs_triples_fn draw, void* user,
                                int32_t* keep_idx, int64_t* n_keep, double* err,
                                double* inlier_ratio);

/* ---------------------------------------------------------------- point-set utilities ------ */
/* v = s R p + t ; n = R n. [ref: Alignment/Alignment.cpp:31-34, Processor/Processor.cpp:1025-1026]
 * normals may be NULL. In place on host doubles. */
int mvs_apply_srt(mvs_ctx* ctx, double* pts_xyz, double* normals_xyz, int64_t n, double scale,
                  const double R[9], const double t[3]);
/* p_k = (1/s) R^T (p - t); n_k = R^T n. [ref: Processor/Processor.cpp:1177-1189] */
int mvs_apply_srt_inverse(mvs_ctx* ctx, double* pts_xyz, double* normals_xyz, int64_t n,
                          double scale, const double R[9], const double t[3]);
/* Device-resident float4 variant (32 B/point streamed). */
int mvs_apply_srt_dev_f32(mvs_ctx* ctx, const float* in_xyzw_dev, float* out_xyzw_dev, int64_t n,
                          double scale, const double R[9], const double t[3]);
/* [ref: SetUtils/PointSetUtils.cpp:3-62] barycenter, bbox (min,max), PCA pivots: columns of
 * `pivots` (row-major 3x3) are unit eigenvectors of the covariance by descending eigenvalue;
 * sign convention: largest-magnitude component positive (the reference's sign is
 * solver-dependent). Any output may be NULL. */
int mvs_pointset_stats(mvs_ctx* ctx, const double* pts_xyz, int64_t n, double barycenter[3],
                       double bbox_min[3], double bbox_max[3], double pivots[9],
                       double eigenvalues[3]);
/* min/max over points of axis.(p - origin)/|axis|^2, optionally restricted to labels[i]==label
 * (labels NULL = all). [ref: Alignment/Alignment.cpp:285-295, :506-528] */
int mvs_project_range(mvs_ctx* ctx, const double* pts_xyz, int64_t n, const int32_t* labels,
                      int32_t label, const double origin[3], const double axis[3], double* tmin,
                      double* tmax, int64_t* argmin, int64_t* argmax);

/* ---------------------------------------------------------------- alignment ---------------- */
typedef struct mvs_align_params {
  double dist_thres;    /* ParamParser::dist_thres  [ref: Parameter/ParamParser.cpp:5-43] */
  double ground_ratio;  /* 0.28 [ref: Alignment/Alignment.cpp:187] */
  int32_t remove_ground; /* 1 = run RemoveGround first (reference behaviour) */
  int32_t local_align;   /* 1 = run the per-limb LocalAlignment stage */
  int32_t neck_label;    /* PART enum value of NECK [ref: PartRecognition/PartRecognition.h:13-30] */
  int32_t reserved;
} mvs_align_params;
int mvs_align_default_params(mvs_align_params* p);
/* [ref: Alignment/Alignment.h:21-28 Align] src (template) is transformed in place; tgt has its
 * ground removed and is compacted (n_tgt, n_tgt_facets updated); facets re-indexed.
 * src_labels: per-template-vertex part labels (the reference loads them from Template/part/parts
 * inside Align, Alignment.cpp:38-41 — file I/O stays with the caller). out_tgt_labels (capacity
 * = input *n_tgt) receives the transferred labels (may be NULL). */
int mvs_align(mvs_ctx* ctx, double* src_xyz, double* src_normals, int64_t n_src,
              const int32_t* src_labels, double* tgt_xyz, double* tgt_normals, int64_t* n_tgt,
              int32_t* tgt_facets, int64_t* n_tgt_facets, const double view_ray[3],
              const mvs_align_params* params, int32_t* out_tgt_labels, double ground_ray[3]);
/* [ref: Alignment/Alignment.cpp:79-233 RemoveGround] */
int mvs_remove_ground(mvs_ctx* ctx, double* pts_xyz, double* normals_xyz, int64_t* n,
                      int32_t* facets, int64_t* n_facets, const mvs_align_params* params,
                      double ground_ray[3]);
/* [ref: Alignment/Alignment.cpp:618-654 RetainConnectRegion + SetUtils/UnionSetUtils.cpp] host
 * union-find (sequential, stays on the host by design — SURVEY §8(a) A7). */
int mvs_retain_connected(double* pts_xyz, double* normals_xyz, int64_t* n, int32_t* facets,
                         int64_t* n_facets);

/* ---------------------------------------------------------------- deformation -------------- */
typedef struct mvs_gn_params {
  int32_t gn_iters;      /* Gauss-Newton iterations (C3: 10) */
  int32_t pcg_iters;     /* PCG iterations per GN step (C3: 20, fixed) */
  int32_t corr_mode;     /* 0: node targets from the reference correspondence rules
                               (Deformation.cpp:266-381) as soft constraints on the node vertices
                            1: all-vertex exact-NN correspondences re-derived every GN iteration
                            2: caller-provided per-vertex targets (mvs_deform_set_targets) */
  int32_t outer_rounds;  /* reference `counter` (Deformation.cpp:252) — 1 */
  double w_rot, w_reg, w_fit; /* 1 / 10 / 100 (Sumner et al. 2007) */
  double lambda;         /* Levenberg damping added to the diagonal, 1e-6 */
  double pcg_tol;        /* relative residual early stop; 0 = run exactly pcg_iters */
  double max_corr_dist2; /* corr_mode 1: reject pairs with d2 > this (<=0: off) */
  double min_corr_cos;   /* corr_mode 1: reject if n_src.n_tgt < this (<= -1: off) */
} mvs_gn_params;
int mvs_gn_default_params(mvs_gn_params* p);

/* [ref: Deformation/Deformation.cpp:29-46 ctor] points/normals n x 3 doubles (cast to float32 as
 * the reference does for every FLANN query), facets 3 x n_facets ints (0-based). */
int mvs_deform_create(mvs_ctx* ctx, const double* pts_xyz, const double* normals_xyz, int64_t n,
                      const int32_t* facets, int64_t n_facets, mvs_deform** out);
int mvs_deform_create_f32(mvs_ctx* ctx, const float* pts_xyz, const float* normals_xyz, int64_t n,
                          const int32_t* facets, int64_t n_facets, mvs_deform** out);
int mvs_deform_destroy(mvs_deform* h);
/* [ref: Deformation/Deformation.cpp:63-106 UniformSampling] sequential greedy: scan vertices in
 * index order, keep i, suppress its `suppress_k` (16) nearest. */
int mvs_deform_uniform_sampling(mvs_deform* h, int suppress_k);
/* Deterministic farthest-point sampling to a prescribed node count (C3: 4096, C5: 20000). */
int mvs_deform_sample_fps(mvs_deform* h, int64_t n_nodes, int64_t first_vertex);
int mvs_deform_set_nodes(mvs_deform* h, const int32_t* vertex_ids, int64_t n_nodes);
int mvs_deform_get_nodes(mvs_deform* h, int32_t* vertex_ids, int64_t capacity, int64_t* n_nodes);
/* Build per-vertex node weights (k nearest + 1), the node graph
 * [ref: Deformation/Deformation.cpp:108-153 KNearestNeighbor(K)] and the block-sparse pattern.
 * Called implicitly by mvs_deform_run when needed. */
int mvs_deform_build_graph(mvs_deform* h, int node_k, int graph_k);
int mvs_deform_get_weights(mvs_deform* h, int32_t* node_idx /*n x 4*/, float* weights /*n x 4*/);
int mvs_deform_get_graph(mvs_deform* h, int32_t* nbr_idx /*m x (graph_k+1)*/);
/* Target scan. [ref: Deformation/Deformation.cpp:238-246] */
int mvs_deform_set_target(mvs_deform* h, const double* tpts_xyz, const double* tnormals_xyz,
                          int64_t n_target);
int mvs_deform_set_target_f32(mvs_deform* h, const float* tpts_xyz, const float* tnormals_xyz,
                              int64_t n_target);
/* Target = the valid pixels of an inverse-depth image, unprojected and compacted in pixel order on
 * the device: mvs_unproject_depth [ref: Image3D/Image3D.cpp:92-106] + "keep the valid points"
 * [ref: Processor/Processor.cpp:1122-1137] + mvs_deform_set_target without the trip through host
 * vectors (same points, same order, same float32 rounding).  n_target (optional): valid count. */
int mvs_deform_set_target_depth(mvs_deform* h, const float* inv_depth, const mvs_camera* cam,
                                double min_dsp, double max_dsp, int64_t* n_target);
int mvs_deform_set_target_dev_f32(mvs_deform* h, const float* tpts_xyzw_dev,
                                  const float* tnormals_xyzw_dev, int64_t n_target);
/* corr_mode 2 targets: per-vertex xyz + weight (0 = unconstrained). */
int mvs_deform_set_targets(mvs_deform* h, const double* targets_xyz, const float* weights,
                           int64_t n);
/* Node correspondence stage alone. [ref: Deformation/Deformation.cpp:266-357 and :359-381]
 * controls: m x 3 smoothed target positions, valid: m flags (before smoothing). */
int mvs_deform_node_targets(mvs_deform* h, double proj_len_err, double proj_dist_err,
                            int smooth_sweeps, double* controls_xyz, int32_t* valid);
/* [ref: Deformation/Deformation.h:241-245 Deform(tpts,tnormals,projLenErr,projDistErr)] — target
 * set with mvs_deform_set_target*; runs the correspondence stage and the embedded-deformation
 * Gauss-Newton solve (replaces CGAL Surface_mesh_deformation, Deformation.cpp:383-400). */
int mvs_deform_run(mvs_deform* h, double proj_len_err, double proj_dist_err,
                   const mvs_gn_params* params);
/* [ref: Deformation/Deformation.h:237-240 Deform(tpts,tnormals,threshold)] */
int mvs_deform_run_threshold(mvs_deform* h, double threshold, const mvs_gn_params* params);
/* The correspondence stage of that variant alone [ref: Deformation/Deformation.cpp:186-203]: per
 * node the 1-NN target point, valid iff d2 <= threshold and cos(dir, node normal) >= 0.5; invalid
 * nodes keep their own position.  controls_xyz m x 3, valid m (either may be NULL). */
int mvs_deform_node_targets_threshold(mvs_deform* h, double threshold, double* controls_xyz,
                                      int32_t* valid);
/* As-rigid-as-possible mode of Deform(): the solver the reference actually runs, CGAL 4.6
 * Surface_mesh_deformation [ref: Deformation/Deformation.cpp:256-260 insert_roi_vertices(all),
 * :383-392 every sampled node a control vertex, :393 preprocess(), :398 deform(5, 1e-4), :400
 * overwrite_initial_geometry()].  CGAL is not vendored: the published algorithm (Sorkine & Alexa
 * 2007; local rotations + cotangent-Laplacian global solve) is restated, parity with CGAL's own
 * floating point unpinned.  The global solve is Jacobi-preconditioned CG on the device. */
typedef struct mvs_arap_params {
  int32_t iters;     /* local/global iterations, reference: 5 */
  int32_t cg_iters;  /* CG iterations per global step (upper bound when cg_tol > 0) */
  double tol;        /* stop when |E_prev - E| <= tol * E, reference: 1e-4; 0 = run all iters */
  double cg_tol;     /* relative residual stop of the global solve; 0 = run exactly cg_iters */
} mvs_arap_params;
int mvs_arap_default_params(mvs_arap_params* p);
/* control vertices = the sampled nodes (mvs_deform_get_nodes order) moved to ctrl_xyz (m x 3);
 * result readable with mvs_deform_get_vertices/normals; energy / iterations_run may be NULL. */
int mvs_deform_arap(mvs_deform* h, const double* ctrl_xyz, const mvs_arap_params* params,
                    double* energy, int32_t* iterations_run);
/* [ref: Deformation/Deformation.h:241-245 Deform(tpts,tnormals,projLenErr,projDistErr)] node
 * correspondences + smoothing (Deformation.cpp:266-381), then the ARAP solve, then
 * overwrite_initial_geometry (the deformed mesh becomes the rest mesh). */
int mvs_deform_run_arap(mvs_deform* h, double proj_len_err, double proj_dist_err,
                        const mvs_arap_params* params);
/* One Gauss-Newton iteration on device-resident state (bench "step"). */
int mvs_deform_gn_begin(mvs_deform* h, const mvs_gn_params* params);
int mvs_deform_gn_step(mvs_deform* h);
int mvs_deform_gn_end(mvs_deform* h);
/* The whole fixed-iteration solve in ONE call: gn_begin + params->gn_iters x gn_step + gn_end, every
 * launch enqueued from C++ with no host synchronisation in between (the loop Deformation::Deform
 * runs inside `deformMesh.deform(...)`, Deformation/Deformation.cpp:398); one stream synchronise
 * at the end.  Result readable with mvs_deform_get_vertices*. */
int mvs_deform_solve(mvs_deform* h, const mvs_gn_params* params);
/* Correspondences of the LAST Gauss-Newton step (corr_mode 1): out_idx[i] = index of the target
 * point that vertex i (original vertex order) was matched to — the exact nearest neighbour, lowest
 * index on ties, of the vertex deformed by the node state the step started from. -1 before the first
 * step or outside this rank's vertex shard. */
int mvs_deform_get_correspondences(mvs_deform* h, int32_t* out_idx /* n */);
/* energy terms after the last step: fit, rot, reg (weighted sums of squares) */
int mvs_deform_get_energy(mvs_deform* h, double energy[3]);
int mvs_deform_get_node_state(mvs_deform* h, float* A_t /* m x 12: a1,a2,a3 (columns), t */);
int mvs_deform_set_node_state(mvs_deform* h, const float* A_t);
/* [ref: Deformation/Deformation.h:233 Polyhedron() / :174-221 exportOBJ] current (deformed)
 * vertices and recomputed normals (sum of unit facet normals, normalised). */
int mvs_deform_get_vertices(mvs_deform* h, double* pts_xyz);
int mvs_deform_get_normals(mvs_deform* h, double* normals_xyz);
int mvs_deform_get_vertices_f32(mvs_deform* h, float* pts_xyz);
int mvs_deform_get_facets(mvs_deform* h, int32_t* facets, int64_t* n_facets);
int mvs_deform_export_obj(mvs_deform* h, const char* filename);
/* Assembled normal equations of the last GN step (tests / inspection), Kronecker-compressed:
 * block-CSR row_ptr (m+1), col (nnz, ascending, diagonal included), blocks nnz x 16 (4x4 factors
 * B_jl with H_jl = kron(B_jl, I3)), rot m x 81 (dense 9x9 rotation-term blocks), rhs m x 12
 * (J^T r).  Any pointer may be NULL; *nnz receives the block count. */
int mvs_deform_get_system(mvs_deform* h, int32_t* row_ptr, int32_t* col, float* blocks, float* rot,
                          float* rhs, int64_t* nnz);
/* Sizes of the solver structures: stats[0..5] = vertices, nodes, pattern blocks (nnz), runs of
 * equal node 4-tuples in the sorted vertex order, facets, target points. */
int mvs_deform_get_stats(mvs_deform* h, int64_t stats[6]);

/* ---------------------------------------------------------------- Model2Depth -------------- */
/* [ref: Model2Depth/Model2Depth.h:17-54 SetInput/Run, Model2Depth.cpp:93-156] z-buffer render of a
 * triangle mesh into each camera; output ncams x h x w float32 inverse depth, row-major, top-left
 * origin, 0 = empty — the `.raw` wire format (Common/Utils.h:177-185). znear/zfar as
 * Model2Depth.cpp:19-20 (0.01 / 2000) when <= 0 is passed. All cameras share w,h. */
int mvs_model2depth(mvs_ctx* ctx, const float* pts_xyz, int64_t n, const int32_t* facets,
                    int64_t n_facets, const mvs_camera* cams, int ncams, float znear, float zfar,
                    float* inv_depth);
int mvs_model2depth_f64(mvs_ctx* ctx, const double* pts_xyz, int64_t n, const int32_t* facets,
                        int64_t n_facets, const mvs_camera* cams, int ncams, float znear,
                        float zfar, float* inv_depth);
int mvs_model2depth_dev(mvs_ctx* ctx, const float* pts_xyzw_dev, int64_t n,
                        const int32_t* facets_dev, int64_t n_facets, const mvs_camera* cams,
                        int ncams, float znear, float zfar, float* inv_depth_dev);
/* [ref: Common/Utils.h:166-185 LoadDepth/SaveDepth] */
int mvs_save_depth_raw(const char* filename, const float* inv_depth, int w, int h);
int mvs_load_depth_raw(const char* filename, float* inv_depth, int w, int h);
/* [ref: Depth2Model/Depth2Model.cpp:7-82 SaveModel] inverse-depth image -> triangle mesh on the
 * pixel grid: a vertex per pixel with min_dsp <= d <= max_dsp (row-major order), two triangles per
 * quad when all three depth differences are <= smooth_threshold * (max_dsp - min_dsp) / 100.
 * out_xyz capacity w*h*3 doubles, tex_index capacity w*h (pixel id of each vertex), facets capacity
 * 6*(w-1)*(h-1) ints; any of them may be NULL to query the counts only. */
int mvs_depth2model(mvs_ctx* ctx, const float* inv_depth, const mvs_camera* cam, double min_dsp,
                    double max_dsp, double smooth_threshold, double* out_xyz, int32_t* tex_index,
                    int32_t* facets, int64_t* n_points, int64_t* n_facets);
/* [ref: Processor/Processor.cpp:66-115 CheckConsistencyCore] zeroes every pixel of `depth` (w*h
 * inverse depths, in place) that is outside [min_dsp, max_dsp] or whose 3-D point does not
 * re-project within reproj_err pixels through every reference frame.  ref_depths: nref x w*h. */
int mvs_check_consistency(mvs_ctx* ctx, float* depth, const mvs_camera* cam, const float* ref_depths,
                          const mvs_camera* ref_cams, int nref, double min_dsp, double max_dsp,
                          double reproj_err);
/* [ref: Image3D/Image3D.cpp:92-106 SolveUnProjectionD] inverse-depth image -> world points;
 * valid iff min_dsp <= d <= max_dsp. out_xyz: w*h*3 doubles, valid: w*h flags. */
int mvs_unproject_depth(mvs_ctx* ctx, const float* inv_depth, const mvs_camera* cam,
                        double min_dsp, double max_dsp, double* out_xyz, int32_t* valid);
/* The same unprojection [ref: Camera/Camera.cpp:40-44,61-67 GetWorldCoordFromImg] for a list of
 * pixels uv (n x 2: column u, row v) — the matched pixels Processor feeds to SRTSolver
 * [ref: Processor/Processor.cpp:806-811] — bit-identical to the rows of mvs_unproject_depth. */
int mvs_unproject_pixels(mvs_ctx* ctx, const float* inv_depth, const mvs_camera* cam, double min_dsp,
                         double max_dsp, const int32_t* uv, int64_t n, double* out_xyz,
                         int32_t* valid);

/* ---------------------------------------------------------------- view pairs ---------------- */
/* The hot path of a list of INDEPENDENT view pairs in one call — what Processor runs per pair
 * [ref: Processor/Processor.cpp:629-826 `for k` pair loop; per pair :751-765 RemoveOutliers,
 * :806-823 SRTSolver::EstimateTransform + ResidualError + chain, :1025-1026 apply, :1130-1137
 * Deformation, :1190-1191 Model2Depth]:
 *   matched pixels -> unprojected matches -> RemoveOutliers -> closed-form SRT + residual ->
 *   Depth2Model mesh of view A moved by (s, R, t) -> deformation-graph Gauss-Newton onto the valid
 *   points of view B -> Model2Depth of the deformed mesh into camera B.
 * Inputs are caller-owned HOST buffers (the reference reads the depth maps from disk); optional
 * outputs are caller-owned host buffers that may be NULL.  The pairs are worked through by
 * `workers` host threads of the library, each with its own stream, so one pair's host glue
 * overlaps the others' kernels; results do not depend on the worker count. */
typedef struct mvs_pair_job {
  const float* depth_a;      /* w*h inverse depths of view A (the template view) */
  const mvs_camera* cam_a;
  const float* depth_b;      /* w*h inverse depths of view B (the target view) */
  const mvs_camera* cam_b;
  const int32_t* matches;    /* n_matches x 4 pixel pairs (u_a, v_a, u_b, v_b) */
  int64_t n_matches;
  uint64_t seed;             /* RANSAC sample seed of this pair */
  /* optional outputs (NULL = not wanted) */
  int32_t* kept_idx;         /* capacity n_matches: indices of the surviving matches */
  double* template_xyz;      /* capacity w*h*3: view-A mesh vertices in the frame of view B */
  int32_t* facets;           /* capacity 6*(w-1)*(h-1) */
  double* deformed_xyz;      /* capacity w*h*3: deformed vertices */
  float* render;             /* w*h: Model2Depth of the deformed mesh in camera B */
} mvs_pair_job;
typedef struct mvs_pair_params {
  double min_dsp, max_dsp;       /* config.txt:23-24 (0.0025, 0.3) */
  double smooth_threshold;       /* config.txt:38 (0.12) */
  mvs_outlier_params outliers;   /* seed is taken from the job */
  int64_t n_nodes;               /* deformation-graph nodes (farthest point sampling), 512 */
  mvs_gn_params gn;              /* corr_mode 1, 4 GN x 20 PCG by default */
  int32_t workers;               /* host threads; 0 = min(8, hardware threads / 2).  Each worker owns a
                                    stream and a device arena for the temporaries of a pair */
  int32_t reserved;
} mvs_pair_params;
typedef struct mvs_pair_result {
  int32_t status;                /* MVS_OK or the error of this pair */
  int32_t reserved;
  double scale, R[9], t[3];      /* x_B = s R x_A + t */
  double err;                    /* RemoveOutliers' err_all / size of the last round */
  double residual;               /* SRTSolver::ResidualError at (s, R, t) */
  int64_t n_kept, n_template, n_facets, n_target;
} mvs_pair_result;
int mvs_pair_default_params(mvs_pair_params* p);
int mvs_stitch_pairs(mvs_ctx* ctx, const mvs_pair_job* jobs, int64_t n_jobs,
                     const mvs_pair_params* params, mvs_pair_result* results);

/* ---------------------------------------------------------------- multi-GPU ---------------- */
/* One process per GPU. id: 128-byte NCCL unique id produced on rank 0 and broadcast by the
 * caller's launcher (torch.distributed / MPI / file). */
int mvs_comm_unique_id(char id[128]);
int mvs_comm_init(mvs_ctx* ctx, int nranks, int rank, const char id[128]);
int mvs_comm_destroy(mvs_ctx* ctx);
/* Vertex-sharded GN: this rank owns vertices [v_begin, v_end) of the mesh (positions in the solver's
 * sorted vertex order) for the fit term; node state is replicated; the partial J^T J / J^T r blocks
 * are all-reduced over NCCL as 64-bit fixed-point integers.  Each boundary is moved down to the start
 * of the assembly group (4 runs of equal node tuples) that contains it, so ranks that split [0, n) at
 * common points tile it exactly and the all-reduced system is BITWISE the unsharded one. */
int mvs_deform_set_shard(mvs_deform* h, int64_t v_begin, int64_t v_end);
/* On a context with a communicator the partial systems are summed over NVLink peer memory (each
 * rank's system lives in a block the other ranks map through CUDA IPC; ncclAllReduce is the fallback,
 * also forced with MVS_NO_PEER=1).  Build the graph AFTER mvs_comm_init, run the same sequence of
 * solves on every rank, and destroy sharded handles on all ranks together (after a barrier of the
 * launcher): a rank must not free a block its peers still map. */
/* Interleaved sharding: [0, n) is cut into nranks x chunks_per_rank chunks (boundaries snapped like
 * set_shard) and this rank owns chunks rank, rank + nranks, ...  The correspondence pass costs more
 * where the template is far from the scan, so contiguous ranges do not balance the ranks; a few
 * chunks per rank spread over the mesh do.  Same bits as the unsharded solve.  nranks = 0 returns to
 * the plain range of mvs_deform_set_shard. */
int mvs_deform_set_shard_interleaved(mvs_deform* h, int rank, int nranks, int chunks_per_rank);
/* Work through this rank's vertex range as `parts` consecutive sub-ranges (group-aligned): the
 * partial sums `parts` ranks would produce, added by the same commutative integer additions the
 * all-reduce performs — the sharded arithmetic checked on ONE GPU (result bitwise equal to
 * parts = 1).  Default 1. */
int mvs_deform_set_shard_parts(mvs_deform* h, int parts);
/* Static partition of `n_units` independent units (view pairs) over ranks: unit u -> u % nranks.
 * [ref: Processor/Processor.cpp:629-826 independent `for k` pair loop] */
int mvs_partition_units(int64_t n_units, int nranks, int rank, int64_t* begin_stride_count /*3*/);

#ifdef __cplusplus
}
#endif
#endif /* MVS_B200_H */
