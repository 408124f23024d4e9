"""ctypes mirror of include/b200pg.h (the C-ABI boundary). Field order must match the header."""
import ctypes as C

c_float_p = C.POINTER(C.c_float)
c_u32_p = C.POINTER(C.c_uint32)

SHAPE_RECTANGLE, SHAPE_TRIMESH = 0, 1
BSDF_DIFFUSE, BSDF_DIELECTRIC, BSDF_ROUGHCONDUCTOR, BSDF_ROUGHPLASTIC, BSDF_NULL = range(5)
DISTR_BECKMANN, DISTR_GGX = 0, 1
PHASE_ISOTROPIC, PHASE_HG = 0, 1
MEDIUM_WOODCOCK, MEDIUM_SIMPSON = 0, 1


class Shape(C.Structure):
    _fields_ = [
        ("type", C.c_int32),
        ("to_world", C.c_float * 16),
        ("bsdf", C.c_int32),
        ("emitter", C.c_int32),
        ("interior_medium", C.c_int32),
        ("exterior_medium", C.c_int32),
        ("n_vertices", C.c_uint32),
        ("n_triangles", C.c_uint32),
        ("positions", c_float_p),
        ("normals", c_float_p),
        ("texcoords", c_float_p),
        ("indices", c_u32_p),
    ]


class Bsdf(C.Structure):
    _fields_ = [
        ("type", C.c_int32),
        ("twosided", C.c_int32),
        ("reflectance", C.c_float * 3),
        ("specular_reflectance", C.c_float * 3),
        ("specular_transmittance", C.c_float * 3),
        ("int_ior", C.c_float),
        ("ext_ior", C.c_float),
        ("eta", C.c_float * 3),
        ("k", C.c_float * 3),
        ("distribution", C.c_int32),
        ("alpha_u", C.c_float),
        ("alpha_v", C.c_float),
        ("sample_visible", C.c_int32),
        ("nonlinear", C.c_int32),
        ("rt_ext_trans", C.c_float * 100),
        ("rt_ext_diff", C.c_float),
        ("rt_int_diff", C.c_float),
    ]


class Emitter(C.Structure):
    _fields_ = [("radiance", C.c_float * 3), ("sampling_weight", C.c_float), ("shape", C.c_int32)]


class Medium(C.Structure):
    _fields_ = [
        ("method", C.c_int32),
        ("scale", C.c_float),
        ("albedo", C.c_float * 3),
        ("phase_type", C.c_int32),
        ("phase_g", C.c_float),
        ("res", C.c_int32 * 3),
        ("aabb_min", C.c_float * 3),
        ("aabb_max", C.c_float * 3),
        ("to_world", C.c_float * 16),
        ("density", c_float_p),
        ("step_size_multiplier", C.c_float),
    ]


class Sensor(C.Structure):
    _fields_ = [
        ("to_world", C.c_float * 16),
        ("fov", C.c_float),
        ("fov_axis", C.c_int32),
        ("near_clip", C.c_float),
        ("far_clip", C.c_float),
        ("medium", C.c_int32),
    ]


class Film(C.Structure):
    _fields_ = [("width", C.c_int32), ("height", C.c_int32), ("filter_stddev", C.c_float),
                ("file_format", C.c_int32), ("component_format", C.c_int32)]


class SceneDesc(C.Structure):
    _fields_ = [
        ("n_shapes", C.c_int32),
        ("n_bsdfs", C.c_int32),
        ("n_emitters", C.c_int32),
        ("n_media", C.c_int32),
        ("shapes", C.POINTER(Shape)),
        ("bsdfs", C.POINTER(Bsdf)),
        ("emitters", C.POINTER(Emitter)),
        ("media", C.POINTER(Medium)),
        ("sensor", Sensor),
        ("film", Film),
        ("sample_count", C.c_int32),
        ("seed", C.c_uint64),
    ]


class IntegratorParams(C.Structure):
    _fields_ = [
        ("max_depth", C.c_int32),
        ("rr_depth", C.c_int32),
        ("strict_normals", C.c_int32),
        ("hide_emitters", C.c_int32),
        ("samples_per_progression", C.c_int32),
        ("max_render_time", C.c_int32),
        ("max_component_value", C.c_float),
        ("use_nee", C.c_int32),
        ("volumetric", C.c_int32),
        ("guiding", C.c_int32),
        ("training_progressions", C.c_int32),
        ("guiding_probability", C.c_float),
        ("guide_max_components", C.c_int32),
        ("guide_max_cell_samples", C.c_int32),
        ("guide_train_discard_film", C.c_int32),
        ("guided_distance", C.c_int32),
        ("max_batch_paths", C.c_int32),
    ]


class Stats(C.Structure):
    _fields_ = [
        ("paths", C.c_uint64),
        ("normal_rays", C.c_uint64),
        ("shadow_rays", C.c_uint64),
        ("path_length_sum", C.c_uint64),
        ("kernel_launches", C.c_uint64),
        ("seconds_total", C.c_double),
        ("seconds_trace", C.c_double),
        ("seconds_shade", C.c_double),
        ("seconds_film", C.c_double),
        ("seconds_train", C.c_double),
        ("bvh_nodes_visited", C.c_uint64),
        ("prims_tested", C.c_uint64),
        ("train_samples", C.c_uint64),
        ("guide_cells", C.c_uint32),
        ("progressions_done", C.c_uint32),
    ]


def default_params():
    """Defaults of progressivepath (integrator.cpp:195-230, progressiveintegrator.cpp:296-300)."""
    p = IntegratorParams()
    p.max_depth = -1
    p.rr_depth = 5
    p.strict_normals = 0
    p.hide_emitters = 0
    p.samples_per_progression = 1
    p.max_render_time = 0
    p.max_component_value = float("inf")
    p.use_nee = 1
    p.volumetric = 0
    p.guiding = 0
    p.training_progressions = 0
    p.guiding_probability = 0.5
    p.guide_max_components = 16
    p.guide_max_cell_samples = 32768
    p.guide_train_discard_film = 0
    p.guided_distance = 0
    p.max_batch_paths = 0
    return p
