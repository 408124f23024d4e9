/*
 * b200pg.h -- C-ABI boundary of the B200-native guided path tracer.
 *
 * This header is the drop-in boundary for the reference's integrator plugin
 * interface (C++-ABI there):
 *
 *   reference                                              here
 *   ---------------------------------------------------    --------------------------
 *   MTS_EXPORT_PLUGIN / CreateInstance(const Properties&)  b200pg_integrator_create
 *     include/mitsuba/core/cobject.h:99-107
 *   Integrator::preprocess / render / postprocess          b200pg_render
 *     include/mitsuba/render/integrator.h:61-107
 *   Integrator::cancel (async)                             b200pg_cancel
 *     include/mitsuba/render/integrator.h:86
 *   Film::put / Film::develop / Film::getStorage           b200pg_film_read / _write
 *     include/mitsuba/render/film.h:51, src/films/hdrfilm.cpp:391-546
 *   SceneHandler (XML -> objects)                          b200pg_scene_load_xml
 *     src/librender/scenehandler.cpp:197-760
 *   Statistics (rays traced, avg. path length)             b200pg_stats
 *     src/librender/skdtree.cpp:46-47, progressive_path.cpp:26
 *   ProgressiveMonteCarloIntegrator pre/postprogression    b200pg_progression_* (training loop hooks)
 *     include/mitsuba/render/progressiveintegrator.h:11-86
 *
 * Plain pointers and sizes only; no C++ or torch types cross this boundary.
 * All functions return 0 on success and a negative code on error unless noted;
 * the message is available from b200pg_last_error() (thread-local).
 */
#ifndef B200PG_H
#define B200PG_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200PG_VERSION 100 /* 0.1.0 */

/* ------------------------------------------------------------------ */
/*  Scene description (flat POD). Field names follow the XML names of  */
/*  the reference plugins wherever one exists.                         */
/* ------------------------------------------------------------------ */

enum B200pgShapeType { B200PG_SHAPE_RECTANGLE = 0, B200PG_SHAPE_TRIMESH = 1 };

enum B200pgBsdfType {
    B200PG_BSDF_DIFFUSE = 0,        /* src/bsdfs/diffuse.cpp        */
    B200PG_BSDF_DIELECTRIC = 1,     /* src/bsdfs/dielectric.cpp     */
    B200PG_BSDF_ROUGHCONDUCTOR = 2, /* src/bsdfs/roughconductor.cpp */
    B200PG_BSDF_ROUGHPLASTIC = 3,   /* src/bsdfs/roughplastic.cpp   */
    B200PG_BSDF_NULL = 4            /* src/bsdfs/null.cpp           */
};

enum B200pgDistribution { B200PG_DISTR_BECKMANN = 0, B200PG_DISTR_GGX = 1 };

enum B200pgPhaseType { B200PG_PHASE_ISOTROPIC = 0, B200PG_PHASE_HG = 1 };

enum B200pgMediumMethod { B200PG_MEDIUM_WOODCOCK = 0, B200PG_MEDIUM_SIMPSON = 1 };

/* 4x4 row-major matrices everywhere (Matrix4x4 m[row][col], transform.h). */

typedef struct B200pgShape {
    int32_t type;            /* B200pgShapeType */
    float to_world[16];      /* rectangle: objectToWorld incl. flipNormals (rectangle.cpp:79-84).
                                trimesh: identity (vertices are already in world space). */
    int32_t bsdf;            /* index into bsdfs, -1 = default per Shape::configure (shape.cpp:48-70) */
    int32_t emitter;         /* index into emitters, -1 = none */
    int32_t interior_medium; /* index into media, -1 = none */
    int32_t exterior_medium;
    uint32_t n_vertices, n_triangles; /* trimesh only */
    const float *positions;           /* 3*n_vertices */
    const float *normals;             /* 3*n_vertices or NULL (face normals) */
    const float *texcoords;           /* 2*n_vertices or NULL */
    const uint32_t *indices;          /* 3*n_triangles */
} B200pgShape;

typedef struct B200pgBsdf {
    int32_t type;     /* B200pgBsdfType */
    int32_t twosided; /* wrapped in a `twosided` adapter (twosided.cpp:117-195), same BRDF on both sides */
    float reflectance[3];            /* diffuse `reflectance` / roughplastic `diffuseReflectance` */
    float specular_reflectance[3];   /* `specularReflectance` */
    float specular_transmittance[3]; /* `specularTransmittance` */
    float int_ior, ext_ior;          /* dielectric / roughplastic */
    float eta[3], k[3];              /* roughconductor (already divided by extEta) */
    int32_t distribution;            /* B200pgDistribution */
    float alpha_u, alpha_v;
    int32_t sample_visible;          /* must be 1 (reference default, microfacet.h:138) */
    int32_t nonlinear;               /* roughplastic */
    /* roughplastic tabulated rough transmittance, reduced as rtrans.h:292-388 does
       (filled by the scene loader; 100-entry 1-D table + scalars) */
    float rt_ext_trans[100];
    float rt_ext_diff;               /* external evalDiffuse (alpha, eta fixed) */
    float rt_int_diff;               /* internal (1/eta) evalDiffuse(alpha) */
} B200pgBsdf;

typedef struct B200pgEmitter { /* `area`, src/emitters/area.cpp */
    float radiance[3];
    float sampling_weight;
    int32_t shape; /* owning shape */
} B200pgEmitter;

typedef struct B200pgMedium { /* `heterogeneous` + `gridvolume`, src/medium/heterogeneous.cpp */
    int32_t method;          /* B200pgMediumMethod */
    float scale;             /* `scale` */
    float albedo[3];         /* constvolume albedo */
    int32_t phase_type;      /* B200pgPhaseType */
    float phase_g;           /* hg `g` */
    int32_t res[3];          /* grid resolution nx, ny, nz */
    float aabb_min[3], aabb_max[3];
    float to_world[16];      /* volume toWorld (identity supported) */
    const float *density;    /* nx*ny*nz float32, x fastest (gridvolume.cpp:56-89) */
    float step_size_multiplier; /* `stepSize` factor for simpson, 0 = auto */
} B200pgMedium;

typedef struct B200pgSensor { /* `perspective`, src/sensors/perspective.cpp */
    float to_world[16];
    float fov;          /* degrees */
    int32_t fov_axis;   /* 0 = x, 1 = y, 2 = diagonal, 3 = smaller, 4 = larger */
    float near_clip, far_clip;
    int32_t medium;     /* index or -1: medium the camera sits in */
} B200pgSensor;

typedef struct B200pgFilm { /* `hdrfilm` + `gaussian` rfilter */
    int32_t width, height;
    float filter_stddev; /* gaussian stddev, radius = 4*stddev (gaussian.cpp:33-37) */
    int32_t file_format;      /* hdrfilm `fileFormat`: 0 = openexr (default, hdrfilm.cpp:212-226), 1 = pfm, 2 = rgbe */
    int32_t component_format; /* hdrfilm `componentFormat`: 0 = float16 (default, :219-220), 1 = float32 */
} B200pgFilm;

typedef struct B200pgSceneDesc {
    int32_t n_shapes, n_bsdfs, n_emitters, n_media;
    const B200pgShape *shapes;
    const B200pgBsdf *bsdfs;
    const B200pgEmitter *emitters;
    const B200pgMedium *media;
    B200pgSensor sensor;
    B200pgFilm film;
    int32_t sample_count; /* sampler `sampleCount` */
    uint64_t seed;
} B200pgSceneDesc;

/* Integrator parameters. Names/defaults mirror the XML parameters of
 * `progressivepath` / `progressivevolpath` (integrator.cpp:195-230,
 * progressiveintegrator.cpp:296-300, progressive_path.cpp:117).
 * Guiding parameters are this repo's own (the guided plugin is not in the
 * reference snapshot). */
typedef struct B200pgIntegratorParams {
    int32_t max_depth;               /* maxDepth, -1 = infinite */
    int32_t rr_depth;                /* rrDepth, default 5 */
    int32_t strict_normals;          /* strictNormals */
    int32_t hide_emitters;           /* hideEmitters */
    int32_t samples_per_progression; /* samplesPerProgression, default 1 */
    int32_t max_render_time;         /* maxRenderTime seconds, 0 = sample budget */
    float max_component_value;       /* maxComponentValue, inf = no clamp */
    int32_t use_nee;                 /* useNee */
    int32_t volumetric;              /* 0 = progressivepath, 1 = progressivevolpath */
    /* guiding */
    int32_t guiding;                 /* 0 = off */
    int32_t training_progressions;   /* number of progressions that train the field. On N devices a pass renders N sample blocks
                                      * and one update refits from all of them: under a sample budget ceil(n / N) passes train (the
                                      * single-device number of training samples), under a time budget n passes (n updates) */
    float guiding_probability;       /* one-sample MIS selection probability (default .5) */
    int32_t guide_max_components;    /* K <= 32 */
    int32_t guide_max_cell_samples;  /* spatial split threshold */
    int32_t guide_train_discard_film;/* 1 = training progressions do not contribute to the film */
    int32_t guided_distance;         /* guided free-flight sampling in media */
    int32_t max_batch_paths;         /* wavefront batch size (0 = auto) */
} B200pgIntegratorParams;

typedef struct B200pgStats {
    uint64_t paths;          /* camera samples completed */
    uint64_t normal_rays;    /* closest-hit queries ("Normal rays traced", skdtree.cpp:46) */
    uint64_t shadow_rays;    /* any-hit queries ("Shadow rays traced", skdtree.cpp:47) */
    uint64_t path_length_sum;/* sum of rRec.depth at termination (progressive_path.cpp:310) */
    uint64_t kernel_launches;
    double seconds_total;
    double seconds_trace;    /* CUDA-event time in trace kernels */
    double seconds_shade;
    double seconds_film;
    double seconds_train;
    uint64_t bvh_nodes_visited; /* only filled by the counting trace variant */
    uint64_t prims_tested;
    uint64_t train_samples;
    uint32_t guide_cells;
    uint32_t progressions_done;
} B200pgStats;

/* ------------------------------------------------------------------ */
/*  Entry points                                                       */
/* ------------------------------------------------------------------ */

int b200pg_version(void);
const char *b200pg_last_error(void);

void b200pg_integrator_params_default(B200pgIntegratorParams *p);

/* Scene: either parsed from Mitsuba 0.6 XML (subset, scenehandler.cpp semantics)
 * or handed over as flat arrays. The loader copies everything it needs. */
void *b200pg_scene_load_xml(const char *path, const char *const *defines /* "k=v", NULL-terminated, may be NULL */,
                            char *err, size_t errlen);
void *b200pg_scene_from_arrays(const B200pgSceneDesc *desc);
/* Borrowed view of the flat description owned by the scene handle (valid until destroy). */
const B200pgSceneDesc *b200pg_scene_desc(void *scene);
/* Integrator parameters found in the XML (defaults if the scene came from arrays). */
int b200pg_scene_integrator_params(void *scene, B200pgIntegratorParams *out);
void b200pg_scene_destroy(void *scene);

void *b200pg_integrator_create(void *scene, const B200pgIntegratorParams *params, int device);
/* Integrator::render (integrator.h:86-95; ProgressiveMonteCarloIntegrator::render, progressiveintegrator.cpp:170-220): blocking,
 * runs all progressions -- sample budget (renderSamples, :65-114) or time budget (renderTime, :117-168) -- with the guiding
 * field trained during the first `training_progressions` passes.
 *   device_count <= 1 or devices == NULL: on the integrator's own device.
 *   device_count  > 1: on all listed devices inside this one call, as the reference registers every worker inside one
 *     render() (mitsuba.cpp:278-327, progressiveintegrator.cpp:84-103). devices[0] must be the integrator's own device; the
 *     library replicates scene and field on the other devices (one host thread each), device r renders sample block
 *     g * device_count + r of global pass g, the per-cell EM statistics are summed over NVLink peer memory inside the M-step
 *     kernel, and the workers' films are added into this integrator's film before the call returns. The sample count is
 *     rounded up to a multiple of device_count * samples_per_progression. Peer access between the devices is required.
 * Clears a previous b200pg_cancel. Returns 0, or < 0 with b200pg_last_error(). */
int b200pg_render(void *integ, int device_count, const int *devices);
int b200pg_cancel(void *integ); /* async-safe */

/* Progression-granular control (what ProgressiveMonteCarloIntegrator::renderSamples does,
 * progressiveintegrator.cpp:65-114). `first_sample`/`n_samples` select the per-pixel sample
 * indices rendered by this call so that sample batches can be split across GPUs;
 * `row_begin`/`row_end` restrict to an image band (tile partition). */
int b200pg_progression_render(void *integ, int first_sample, int n_samples, int row_begin, int row_end);
/* Guiding control for progression-granular use: `record` = the following progressions store path-vertex
 * training samples, `sample` = they draw directions from the field (one-sample MIS with the BSDF). */
int b200pg_guiding_mode(void *integ, int record, int sample);
/* Training update between progressions (the place of postprogression(), progressiveintegrator.h:40-52):
 *   begin       bin the recorded samples per cell (kd-tree lookup + stable radix sort)
 *   accumulate  E-step over the local samples -> per-cell sufficient statistics on the device
 *   stats_buffer exposes that buffer (cells x (4K + 8) floats) for an external sum over ranks (NCCL allreduce)
 *   update      M-step from the (summed) statistics; commit != 0 on the last EM iteration of the update
 *   end         spatial split of over-full cells; identical on every rank given identical statistics */
int b200pg_train_begin(void *integ, uint32_t *n_samples, uint32_t *n_cells);
int b200pg_train_accumulate(void *integ);
int b200pg_train_stats_buffer(void *integ, void **dev_ptr, size_t *n_floats);
int b200pg_train_update(void *integ, int commit);
int b200pg_train_end(void *integ);

/* The whole training update in one call (begin, n_iter x (E-step, [cross-GPU sum,] M-step), end) without host round
 * trips between the EM iterations; n_iter <= 0 uses the integrator's emIterations. With peers connected (below) the
 * per-cell sufficient statistics are summed over all ranks inside the M-step kernel. */
int b200pg_train(void *integ, int n_iter, uint32_t *n_samples, uint32_t *n_cells);

/* Multi-GPU statistics exchange over NVLink peer memory (one process per GPU, replicated scene and field):
 *   b200pg_comm_local_handle  writes the 64-byte CUDA IPC handle of this rank's exchange block
 *   b200pg_comm_connect       maps the blocks of all `world` ranks (`handles` = world x 64 bytes, gathered by the host
 *                             with whatever transport it has: torch.distributed all_gather, MPI, Mitsuba's own
 *                             StreamBackend sched_remote.cpp:33-119)
 * Afterwards b200pg_train sums the statistics of all ranks in rank order (bit-identical on every rank), so the
 * replicated fields stay identical without a broadcast. All ranks must call b200pg_train the same number of times. */
int b200pg_comm_local_handle(void *integ, void *handle64);
int b200pg_comm_connect(void *integ, int rank, int world, const void *handles);

int b200pg_film_clear(void *integ);
/* Denoiser feature buffers (src/librender/denoiser.cpp:138-144, Denoiser::add: per-pixel running means of the sample
 * colour, albedo and normal over the samples that fall into the pixel). Enabled by b200pg_set_option("feature_buffers", 1)
 * before rendering; cleared by b200pg_film_clear. The reference never fills Sample::albedo / normal (nothing in its tree
 * calls the denoiser), so: first intersection of the camera ray; albedo = diffuse reflectance (diffuse, roughplastic),
 * specular reflectance (roughconductor), 1 (dielectric, null); normal = shading normal (world space); both 0 for rays
 * that leave the scene. b200pg_features_read: out = H*W*10 floats {color.rgb, albedo.rgb, normal.xyz, sample count};
 * b200pg_features_write: float32 OpenEXR with the layers color / albedo / normal of Denoiser::saveBuffers
 * (denoiser.cpp:88-112). */
int b200pg_features_read(void *integ, float *out);
int b200pg_features_write(void *integ, const char *path);
/* Multi-GPU film merge, one process per GPU (every GPU keeps a full-size film of its own sample batches; the
 * reference merges its workers' ImageBlocks in Film::put, renderproc.cpp:141-148). b200pg_film_ipc_handle returns the
 * 64-byte CUDA IPC handle of this handle's device film; b200pg_film_add_peers(rank, world, handles = world x 64 bytes)
 * adds the films of all other ranks into this one, read over NVLink in rank order. The caller makes sure the peers have
 * finished rendering and keep their handles alive until the call returns. */
int b200pg_film_ipc_handle(void *integ, void *handle64);
int b200pg_film_add_peers(void *integ, int rank, int world, const void *handles);
int b200pg_film_device_buffer(void *integ, void **dev_ptr, size_t *n_floats); /* H*W*4: R,G,B,weight */
int b200pg_film_read(void *integ, float *rgbaw /* H*W*5: R,G,B,alpha,weight (imageblock.h:131-138) */);
int b200pg_film_develop(void *integ, float *rgb /* H*W*3 = RGB/weight, fmtconv.cpp:978-1005 */);
/* Progressive preview (what mtsgui does between progressions, renderproc.cpp:141-148 + the film's develop): snapshot the film as
 * it is now and copy it to `rgbaw_pinned` (page-locked host memory, H*W*5) on a second stream while rendering continues;
 * b200pg_film_read_wait blocks until the snapshot has arrived. A new async read first waits for the previous one. */
int b200pg_film_read_async(void *integ, float *rgbaw_pinned);
int b200pg_film_read_wait(void *integ);
/* Multi-GPU previews (the reference merges the workers' image blocks in the master's film, renderproc.cpp:141-148): map the
 * films of the other ranks (handles of b200pg_film_ipc_handle, `world` x 64 bytes, own slot ignored). From then on
 * b200pg_film_read_async delivers own film + the peers' films, summed on the device over NVLink in fixed rank order, so a
 * job needs ONE device->host copy per preview. The films themselves are not modified. world <= 1 or handles == NULL unmaps. */
int b200pg_film_peers_connect(void *integ, int rank, int world, const void *handles);
/* Film::develop to a file (hdrfilm.cpp:487-546), format chosen by the extension: .exr (scanline OpenEXR, uncompressed, channels
 * B G R as float16 or float32 per the film's componentFormat), .pfm (float32), .rgbe / .hdr (Radiance RGBE, flat).
 * The reference's banner (`banner=true` draws a logo into the image, :501-511) is never drawn. */
int b200pg_film_write(void *integ, const char *path);
int b200pg_stats(void *integ, B200pgStats *out);
void b200pg_destroy(void *integ);

/* Measurement helpers (no reference counterpart; the reference only logs "Progression[i] took t s",
 * progressiveintegrator.cpp:314-317).
 *   options: "count_traversal" (0/1: counting variant of the trace kernels fills bvh_nodes_visited /
 *            prims_tested), "timing" (0/1: per-stage CUDA events), "sort_bounces" (n: guided surface progressions
 *            shade bounces 1..n through a permutation that groups the queued paths by guiding cell; 0 = queue order;
 *            default 0, or the value of the environment variable
 *            B200PG_SORT_BOUNCES; per-path results do not depend on it), "feature_buffers" (0/1: accumulate the denoiser
 *            feature buffers, see b200pg_features_read).
 *            Scheduling knobs -- they change the order in which paths are shaded and rays are traversed, never a sample's
 *            value (tests/test_gpu_mesh.py::test_mesh_radiance_does_not_depend_on_the_schedule): "lanes" (1..8 concurrent
 *            wavefront sub-batches per progression, default 2), "overlap_shadow" (0/1: shadow stage of bounce b next to the
 *            closest-hit stage of bounce b + 1, default 1), "lane_major" (0/1 enqueue order, default 0), "trace_spec" (bit 0:
 *            persistent speculative traversal for bounce queues, bit 1: for shadow queues, bit 2: for camera rays; default 3),
 *            "partition" (0 = never, 1 = adaptive hit / miss partition of sparse shade queues and the event partition of the
 *            volumetric path, 2 = every bounce >= 1; default 1), "tail_visits" (node-visit budget of a ray before the
 *            warp-cooperative kernel finishes it; 0 = off, -1 = library default: 96 on scenes with >= 64 k BVH nodes),
 *            "splat_tile" (0/1: film accumulation through a per-warp shared-memory tile, default 1).
 *            "split_levels" (1..16, default 1): spatial split levels per training update of the guiding field (one device);
 *            this one changes the trained field, as the oracle's guideTrain(..., splitLevels) does.
 *   stage times: 5 doubles / 5 launch counts = trace(closest), shade, shadow(any-hit), film, train.
 *   scene_upload: re-sends the compiled scene host->device (bench.py's end-to-end leg). */
int b200pg_set_option(void *integ, const char *name, int value);
int b200pg_stage_times(void *integ, double *seconds5, uint64_t *launches5);
int b200pg_scene_upload(void *integ, size_t *bytes);

/* ------------------------------------------------------------------ */
/*  Per-kernel entry points (HOST buffers in/out; copies are done       */
/*  inside). Used by the parity tests and by bench.py's e2e leg.        */
/* ------------------------------------------------------------------ */

/* rays: n * 8 floats (ox,oy,oz,mint, dx,dy,dz,maxt). hits: n * 4 words (t,u,v as float; prim as u32,
 * 0xFFFFFFFF = miss). prim ids are "global primitive ids": shapes in order, rectangle = 1 prim,
 * trimesh = n_triangles prims (same numbering as ShapeKDTree, skdtree.cpp:53-104). */
int b200pg_k_trace(void *integ, const float *rays, size_t n, int shadow, float *hits_tuv, uint32_t *hits_prim);
/* Same with device pointers, returns elapsed kernel ms in *ms (CUDA events). counts (may be NULL): 2 u64
 * (bvh nodes visited, prims tested) filled by the counting variant. */
int b200pg_k_trace_device(void *integ, const void *d_rays, size_t n, int shadow, void *d_hits, float *ms,
                          uint64_t *counts);

/* BSDF eval/pdf/sample in local coordinates. wi, wo: n*3. u: n*2.
 * out_eval n*3, out_pdf n, out_wo n*3, out_weight n*3, out_spdf n, out_flags n (sampled type bits). */
int b200pg_k_bsdf(void *integ, int bsdf_index, const float *wi, const float *wo, const float *u, size_t n,
                  float *out_eval, float *out_pdf, float *out_wo, float *out_weight, float *out_spdf,
                  uint32_t *out_flags);

/* Radiance of n camera samples with explicit (pixel, sample index) pairs: out n*3. Uses the same
 * counter-based RNG as a full render, so results are comparable sample by sample. */
int b200pg_k_radiance(void *integ, const uint32_t *pixel, const uint32_t *sample_index, size_t n, float *out_rgb);

/* Film splat of n samples (pos n*2, rgb n*3) into a cleared film; read back with b200pg_film_read. */
int b200pg_k_film_splat(void *integ, const float *pos, const float *rgb, size_t n);

/* Medium: trilinear density lookups (p n*3 -> out n), see gridvolume.cpp:337-388. */
int b200pg_k_grid_lookup(void *integ, int medium, const float *p, size_t n, float *out);

/* Medium: free-flight sampling and transmittance along n rays (rays n*8 = o.xyz, mint, d.xyz, maxt), replacing
 * HeterogeneousMedium::sampleDistance / evalTransmittance (heterogeneous.cpp:589-663, 546-587; Woodcock branch) and
 * PhaseFunction::sample (hg.cpp:74-95, isotropic.cpp:62-74). Ray i draws from the stream (seed, pixel = i, sample = 0):
 * first the distance, then the transmittance estimate, then one phase-function sample with wi = -d.
 * out_t[i] = sampled distance (inf: left the medium), out_tr[i] = transmittance estimate (0, 0.5 or 1),
 * out_wo (n*3) / out_pdf = sampled scattering direction and its pdf. */
int b200pg_k_medium_sample(void *integ, int medium, const float *rays, size_t n, float *out_t, float *out_tr,
                           float *out_wo, float *out_pdf);

/* Guiding field kernels (vMF mixtures; this repo's own algorithm, oracle-pinned).
 * Field snapshot layout is documented in DESIGN.md. */
/* pos n*3, dir n*3, u n*3 (lobe selection, u1, u2): out_pdf = pdf of dir, out_dir/out_spdf = sampled direction + pdf */
int b200pg_k_vmm_pdf_sample(void *integ, const float *pos, const float *dir, const float *u, size_t n,
                            float *out_pdf, float *out_dir, float *out_spdf, uint32_t *out_cell);
/* Bin n samples (pos n*3) by guiding cell: out_cell n, out_perm n (stable order), out_offsets n_cells+1. */
int b200pg_k_bin_samples(void *integ, const float *pos, size_t n, uint32_t *out_cell, uint32_t *out_perm,
                         uint32_t *out_offsets, uint32_t *n_cells);
/* A complete training update (bin, n_iter x (E, M), split) over externally supplied samples
 * (pos n*3, dir n*3, weight n, pdf n, dist n). n_iter = 0: E-step only, stats_out (cells*(4K+8) floats, may be NULL)
 * receives the sufficient statistics and the field is left untouched. With n_iter > 0 stats_out (if not NULL) receives
 * the statistics of the LAST iteration, cells = the cell count BEFORE this update's split: size it for
 * b200pg_stats().guide_cells * (4K+8) floats (pass NULL when they are not needed). */
int b200pg_k_em_step(void *integ, const float *pos, const float *dir, const float *weight, const float *pdf,
                     const float *dist, size_t n, int n_iter, float *stats_out);
/* Microbenchmark of the statistics exchange alone (SURVEY.md 8d, config C5 "EM allreduce scaling"): n_iter launches of the
 * fused cross-GPU sum + M-step kernel over n_cells synthetic cells (K = the integrator's guide_max_components), after 3
 * warm-up launches; *ms_per_iter = average device time (CUDA events). Every connected rank must make the same call (the
 * kernel contains the cross-GPU barriers). mode 0: as b200pg_train runs it (form chosen by size); 1: this rank's own
 * buffer only (the M-step share); 2: all-read form (every rank reads every peer's buffer); 3: reduce-scatter +
 * all-gather form (every rank sums its slice of the cells and pushes the sums). Requires b200pg_comm_connect (world = 1
 * is allowed). The guiding field itself is not touched. */
int b200pg_k_em_exchange(void *integ, uint32_t n_cells, int n_iter, int mode, float *ms_per_iter);
/* Field snapshot as 32-bit words: header[8] = {'GUID', nNodes, nCells, K, 0...}, nodes[4*nNodes] = {axis (3 = leaf),
 * split, left | cell, 0}, cell headers[8*nCells] = {running sample count, running weight sum, 0...}, lobes[12*nCells*K] =
 * {pi, mu.xyz, kappa, norm, exp(-2 kappa), 0, S, R.xyz}. Pass out = NULL to query the size. */
int b200pg_field_snapshot(void *integ, uint32_t *out, size_t *n_words);
int b200pg_field_load(void *integ, const uint32_t *in, size_t n_words);

#ifdef __cplusplus
}
#endif
#endif /* B200PG_H */
