// ref_harness.cpp -- TEST INFRASTRUCTURE. C-ABI driver around the REFERENCE's own classes, linked against the reference
// libraries that oracle/Makefile.ref compiles from /root/reference (libmitsuba-core / -render / -hw + the plugins on the path).
// Nothing here restates an algorithm: every number that leaves this file is computed by reference code
// (ShapeKDTree::rayIntersect, BSDF::eval / pdf / sample, Scene::sampleEmitterDirect, ProgressiveMIPathTracer::Li,
// ProgressiveMonteCarloIntegrator::render / renderBlock, ImageBlock::put, HeterogeneousMedium, ...). The file only
//   (1) turns a B200pgSceneDesc (include/b200pg.h, the same flat description the product and the oracle port consume) into
//       reference objects through PluginManager::createObject + Properties, the way scenehandler.cpp would from XML, and
//   (2) plugs in a Sampler (ReplaySampler below) that hands out the counter-based PCG32 stream (seed, pixel, sample) which
//       the product and the oracle port use, so that the reference's transport code can be compared SAMPLE BY SAMPLE
//       (ProgressiveMonteCarloIntegrator::preprocess clones the scene's sampler once per pixel and calls generate(pixel),
//       progressiveintegrator.cpp:43-51, and renderBlock advances it once per sample, :282 -- exactly the hooks needed).
// Used by tests/ (the pin of the oracle port, tests/test_ref_*.py), tests/golden/make_upstream.py (fixtures) and the
// `--impl reference` / cpu_baseline legs of bench.py. The product never loads it.
#include <mitsuba/core/plugin.h>
#include <mitsuba/core/statistics.h>
#include <mitsuba/core/fresolver.h>
#include <mitsuba/core/fstream.h>
#include <mitsuba/core/bitmap.h>
#include <mitsuba/core/sched.h>
#include <mitsuba/core/appender.h>
#include <mitsuba/render/scene.h>
#include <mitsuba/render/trimesh.h>
#include <mitsuba/render/skdtree.h>
#include <mitsuba/render/renderjob.h>
#include <mitsuba/render/renderqueue.h>
#include <mitsuba/render/imageblock.h>
#include <mitsuba/render/volume.h>
#include <mitsuba/render/medium.h>
#include <mitsuba/render/phase.h>
#include <mitsuba/render/progressiveintegrator.h>
#include <omp.h>
#include <execinfo.h>
#include <dlfcn.h>
#include <signal.h>
#include <chrono>
#include <cstdio>
#include <map>
#include "b200pg.h"

using namespace mitsuba;

// ---------------------------------------------------------------------------------------------------------------------------
// The shared random stream: PCG32 (XSH-RR), increment from the pixel, state offset from (seed, sample) through a
// splitmix64 finaliser -- the specification in DESIGN.md section 3 ("one sampler per pixel"), restated here so that this file
// depends on nothing under oracle/ but the reference.
// ---------------------------------------------------------------------------------------------------------------------------
namespace {
inline uint64_t mix64(uint64_t z) {
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}
struct Pcg {
    uint64_t state = 0, inc = 1;
    void init(uint64_t seed, uint32_t pixel, uint32_t sample) {
        inc = ((uint64_t)pixel << 1) | 1ULL;
        state = 0;
        next();
        state += mix64(seed + (uint64_t)sample * 0x9E3779B97F4A7C15ULL);
        next();
    }
    uint32_t next() {
        uint64_t old = state;
        state = old * 6364136223846793005ULL + inc;
        uint32_t x = (uint32_t)(((old >> 18u) ^ old) >> 27u), rot = (uint32_t)(old >> 59u);
        return (x >> rot) | (x << ((32u - rot) & 31u));
    }
    float next1D() { return (float)(next() >> 8) * (1.0f / 16777216.0f); }
};
}  // namespace

class ReplaySampler : public Sampler {
public:
    ReplaySampler(uint64_t seed, int filmWidth, size_t sampleCount, size_t firstSample)
        : Sampler(Properties()), m_seed(seed), m_width(filmWidth), m_first(firstSample), m_pixel(0) {
        m_sampleCount = sampleCount;
        m_sampleIndex = firstSample;
        m_rng.init(m_seed, m_pixel, (uint32_t)m_sampleIndex);
    }
    ref<Sampler> clone() { return new ReplaySampler(m_seed, m_width, m_sampleCount, m_first); }
    void generate(const Point2i &pixel) { setStream((uint32_t)(pixel.y * m_width + pixel.x), m_first); }
    void setStream(uint32_t pixel, size_t sample) {
        m_pixel = pixel;
        m_sampleIndex = sample;
        m_rng.init(m_seed, m_pixel, (uint32_t)m_sampleIndex);
    }
    void advance() {
        ++m_sampleIndex;
        m_rng.init(m_seed, m_pixel, (uint32_t)m_sampleIndex);
    }
    void setSampleIndex(size_t i) {
        m_sampleIndex = i;
        m_rng.init(m_seed, m_pixel, (uint32_t)m_sampleIndex);
    }
    Float next1D() { return m_rng.next1D(); }
    Point2 next2D() {
        Float a = m_rng.next1D();
        Float b = m_rng.next1D();
        return Point2(a, b);
    }
    std::string toString() const { return "ReplaySampler[IndependentSampler-compatible counter stream]"; }
    MTS_DECLARE_CLASS()
private:
    uint64_t m_seed;
    int m_width;
    size_t m_first;
    uint32_t m_pixel;
    Pcg m_rng;
};
MTS_IMPLEMENT_CLASS(ReplaySampler, false, Sampler)

// ---------------------------------------------------------------------------------------------------------------------------
// Start-up (the sequence of src/mitsuba/mitsuba.cpp:422-431 minus SHVector / SceneHandler, which this build leaves out)
// ---------------------------------------------------------------------------------------------------------------------------
static bool g_init = false;
static int g_workers = 0;
static void segvTrace(int) {  // REF_HARNESS_VERBOSE only: where did the reference build fall over
    void *frames[64];
    int n = backtrace(frames, 64);
    backtrace_symbols_fd(frames, n, 2);
    _exit(139);
}
static void ensureInit() {
    if (g_init) return;
    if (getenv("REF_HARNESS_VERBOSE")) signal(SIGSEGV, segvTrace);
    Class::staticInitialization();
    Object::staticInitialization();
    PluginManager::staticInitialization();
    Statistics::staticInitialization();
    Thread::staticInitialization();
    Logger::staticInitialization();
    FileStream::staticInitialization();
    Spectrum::staticInitialization();
    Bitmap::staticInitialization();
    Scheduler::staticInitialization();
    Thread::getThread()->getLogger()->setLogLevel(getenv("REF_HARNESS_VERBOSE") ? EInfo : EError);
    ProgressReporter::setEnabled(false);
    g_init = true;
}

static void stopWorkers() {  // at exit: the LocalWorker threads would otherwise keep the process from ending
    if (g_workers) Scheduler::getInstance()->stop();
    g_workers = 0;
}
static void ensureWorkers(int n) {
    static bool registered = false;
    if (!registered) {
        atexit(stopWorkers);
        registered = true;
    }
    Scheduler *sched = Scheduler::getInstance();
    if (n <= 0) n = omp_get_max_threads();
    if (g_workers == n) return;
    if (g_workers != 0) {  // the worker set of a Scheduler is fixed once started: rebuild it
        sched->stop();
        Scheduler::staticShutdown();
        Scheduler::staticInitialization();
        sched = Scheduler::getInstance();
    }
    for (int i = 0; i < n; ++i) sched->registerWorker(new LocalWorker(-1, formatString("wrk%i", i)));
    sched->start();
    g_workers = n;
}

static thread_local std::string g_err;

// ---------------------------------------------------------------------------------------------------------------------------
// Scene construction
// ---------------------------------------------------------------------------------------------------------------------------
struct RefScene {
    ref<Scene> scene;
    ref<Sensor> sensor;
    ref<Film> film;
    ref<ReplaySampler> sampler;
    ref<SamplingIntegrator> integ;  // created per call (parameters vary)
    std::vector<ref<BSDF>> bsdfs;
    std::vector<ref<Medium>> media;
    std::vector<ref<VolumeDataSource>> densities;
    std::vector<ref<Shape>> shapes;
    std::map<const Shape *, int> shapeIndex;
    std::vector<uint32_t> primOffset;
    std::vector<std::string> tmpFiles;
    uint64_t seed = 0;
    int width = 0, height = 0, sampleCount = 1;
    B200pgIntegratorParams lastParams{};
    bool haveInteg = false, pluginInteg = false;
    ~RefScene() {
        for (auto &f : tmpFiles) remove(f.c_str());
    }
};

static Spectrum rgb(const float *c) {
    Spectrum s;
    s.fromLinearRGB(c[0], c[1], c[2]);
    return s;
}
static Transform xform(const float *m) { return Transform(Matrix4x4(m)); }

template <class T> static ref<T> create(const Properties &p) {
    return static_cast<T *>(PluginManager::getInstance()->createObject(MTS_CLASS(T), p));
}

static ref<BSDF> makeBsdf(const B200pgBsdf &b) {
    const char *distr = b.distribution == B200PG_DISTR_GGX ? "ggx" : "beckmann";
    Properties p;
    switch (b.type) {
        case B200PG_BSDF_DIFFUSE:
            p = Properties("diffuse");
            p.setSpectrum("reflectance", rgb(b.reflectance));
            break;
        case B200PG_BSDF_DIELECTRIC:
            p = Properties("dielectric");
            p.setFloat("intIOR", b.int_ior);
            p.setFloat("extIOR", b.ext_ior);
            p.setSpectrum("specularReflectance", rgb(b.specular_reflectance));
            p.setSpectrum("specularTransmittance", rgb(b.specular_transmittance));
            break;
        case B200PG_BSDF_ROUGHCONDUCTOR:
            p = Properties("roughconductor");
            p.setString("material", "none");  // eta / k are given explicitly: no data/ior/*.spd look-up (roughconductor.cpp:196-214)
            p.setString("distribution", distr);
            p.setFloat("alphaU", b.alpha_u);
            p.setFloat("alphaV", b.alpha_v);
            p.setSpectrum("eta", rgb(b.eta));
            p.setSpectrum("k", rgb(b.k));
            p.setFloat("extEta", 1.0f);
            p.setSpectrum("specularReflectance", rgb(b.specular_reflectance));
            p.setBoolean("sampleVisible", b.sample_visible != 0);
            break;
        case B200PG_BSDF_ROUGHPLASTIC:
            p = Properties("roughplastic");
            p.setString("distribution", distr);
            p.setFloat("alpha", b.alpha_u);
            p.setFloat("intIOR", b.int_ior);
            p.setFloat("extIOR", b.ext_ior);
            p.setSpectrum("diffuseReflectance", rgb(b.reflectance));
            p.setSpectrum("specularReflectance", rgb(b.specular_reflectance));
            p.setBoolean("nonlinear", b.nonlinear != 0);
            p.setBoolean("sampleVisible", b.sample_visible != 0);
            break;
        case B200PG_BSDF_NULL:
            p = Properties("null");
            break;
        default:
            throw std::runtime_error("unknown BSDF type");
    }
    ref<BSDF> inner = create<BSDF>(p);
    inner->configure();
    if (!b.twosided) return inner;
    ref<BSDF> outer = create<BSDF>(Properties("twosided"));
    outer->addChild(inner);
    outer->configure();
    return outer;
}

static std::string writeVol(const B200pgMedium &m) {
    char name[] = "/tmp/ref_harness_XXXXXX";
    int fd = mkstemp(name);
    if (fd < 0) throw std::runtime_error("mkstemp failed");
    FILE *f = fdopen(fd, "wb");
    const char hdr[4] = {'V', 'O', 'L', 3};
    int32_t ints[5] = {1, m.res[0], m.res[1], m.res[2], 1};
    fwrite(hdr, 1, 4, f);
    fwrite(ints, 4, 5, f);
    fwrite(m.aabb_min, 4, 3, f);
    fwrite(m.aabb_max, 4, 3, f);
    fwrite(m.density, 4, (size_t)m.res[0] * m.res[1] * m.res[2], f);
    fclose(f);
    return name;
}

static ref<Medium> makeMedium(RefScene &rs, const B200pgMedium &m) {
    Properties pm("heterogeneous");
    pm.setString("method", m.method == B200PG_MEDIUM_SIMPSON ? "simpson" : "woodcock");
    pm.setFloat("scale", m.scale);
    if (m.step_size_multiplier > 0) pm.setFloat("stepSize", m.step_size_multiplier);
    ref<Medium> med = create<Medium>(pm);
    std::string vol = writeVol(m);
    rs.tmpFiles.push_back(vol);
    Properties pd("gridvolume");
    pd.setString("filename", vol);
    pd.setTransform("toWorld", xform(m.to_world));
    ref<VolumeDataSource> dens = create<VolumeDataSource>(pd);
    dens->configure();
    Properties pa("constvolume");
    pa.setSpectrum("value", rgb(m.albedo));
    ref<VolumeDataSource> alb = create<VolumeDataSource>(pa);
    alb->configure();
    Properties pp(m.phase_type == B200PG_PHASE_HG ? "hg" : "isotropic");
    if (m.phase_type == B200PG_PHASE_HG) pp.setFloat("g", m.phase_g);
    ref<PhaseFunction> phase = create<PhaseFunction>(pp);
    phase->configure();
    med->addChild("density", dens);
    med->addChild("albedo", alb);
    med->addChild(phase);
    med->configure();
    rs.densities.push_back(dens);
    return med;
}

static bool g_noSensor = false;  // ref_scene_create_without_sensor: leave the camera to Scene::configure's fallback
static RefScene *buildScene(const B200pgSceneDesc *d) {
    ensureInit();
    std::unique_ptr<RefScene> rs(new RefScene());
    rs->seed = d->seed;
    rs->width = d->film.width;
    rs->height = d->film.height;
    rs->sampleCount = d->sample_count > 0 ? d->sample_count : 1;
    rs->scene = new Scene();

    for (int i = 0; i < d->n_bsdfs; ++i) rs->bsdfs.push_back(makeBsdf(d->bsdfs[i]));
    for (int i = 0; i < d->n_media; ++i) rs->media.push_back(makeMedium(*rs, d->media[i]));

    // sensor <- film <- rfilter, sampler
    Properties pf("hdrfilm");
    pf.setInteger("width", d->film.width);
    pf.setInteger("height", d->film.height);
    pf.setBoolean("banner", false);
    pf.setString("pixelFormat", "rgb");
    ref<Film> film = create<Film>(pf);
    Properties pr("gaussian");
    pr.setFloat("stddev", d->film.filter_stddev);
    ref<ReconstructionFilter> rf = create<ReconstructionFilter>(pr);
    rf->configure();
    film->addChild(rf);
    film->configure();
    Properties ps("perspective");
    ps.setTransform("toWorld", xform(d->sensor.to_world));
    ps.setFloat("fov", d->sensor.fov);
    static const char *axes[] = {"x", "y", "diagonal", "smaller", "larger"};
    ps.setString("fovAxis", axes[d->sensor.fov_axis]);
    ps.setFloat("nearClip", d->sensor.near_clip);
    ps.setFloat("farClip", d->sensor.far_clip);
    ref<Sensor> sensor = create<Sensor>(ps);
    rs->sampler = new ReplaySampler(d->seed, d->film.width, rs->sampleCount, 0);
    sensor->addChild(film);
    sensor->addChild(rs->sampler);
    if (d->sensor.medium >= 0) sensor->addChild(rs->media[d->sensor.medium]);
    sensor->configure();
    rs->sensor = sensor;
    rs->film = film;
    if (!g_noSensor) rs->scene->addChild(sensor);

    uint32_t primOffset = 0;
    std::vector<int> emitterOfShape(d->n_shapes, -1);
    for (int e = 0; e < d->n_emitters; ++e) emitterOfShape[d->emitters[e].shape] = e;
    for (int i = 0; i < d->n_shapes; ++i) {
        const B200pgShape &s = d->shapes[i];
        ref<Shape> shape;
        if (s.type == B200PG_SHAPE_RECTANGLE) {
            Properties p("rectangle");
            p.setTransform("toWorld", xform(s.to_world));
            shape = create<Shape>(p);
            rs->primOffset.push_back(primOffset);
            primOffset += 1;
        } else {
            ref<TriMesh> mesh = new TriMesh(formatString("mesh%i", i), s.n_triangles, s.n_vertices, s.normals != NULL,
                                            s.texcoords != NULL, false, false, s.normals == NULL);
            Point *pos = mesh->getVertexPositions();
            for (uint32_t v = 0; v < s.n_vertices; ++v) pos[v] = Point(s.positions[3 * v], s.positions[3 * v + 1], s.positions[3 * v + 2]);
            if (s.normals) {
                Normal *nn = mesh->getVertexNormals();
                for (uint32_t v = 0; v < s.n_vertices; ++v) nn[v] = Normal(s.normals[3 * v], s.normals[3 * v + 1], s.normals[3 * v + 2]);
            }
            if (s.texcoords) {
                Point2 *tc = mesh->getVertexTexcoords();
                for (uint32_t v = 0; v < s.n_vertices; ++v) tc[v] = Point2(s.texcoords[2 * v], s.texcoords[2 * v + 1]);
            }
            Triangle *tri = mesh->getTriangles();
            for (uint32_t t = 0; t < s.n_triangles; ++t)
                for (int k = 0; k < 3; ++k) tri[t].idx[k] = s.indices[3 * t + k];
            shape = mesh;
            rs->primOffset.push_back(primOffset);
            primOffset += s.n_triangles;
        }
        if (s.bsdf >= 0) shape->addChild(rs->bsdfs[s.bsdf]);
        if (emitterOfShape[i] >= 0) {
            const B200pgEmitter &e = d->emitters[emitterOfShape[i]];
            Properties pe("area");
            pe.setSpectrum("radiance", rgb(e.radiance));
            pe.setFloat("samplingWeight", e.sampling_weight);
            ref<Emitter> em = create<Emitter>(pe);
            shape->addChild(em);
            em->setParent(shape);  // what SceneHandler::endElement does after addChild (scenehandler.cpp): AreaLight keeps its shape
            em->configure();
        }
        if (s.interior_medium >= 0) shape->addChild("interior", rs->media[s.interior_medium]);
        if (s.exterior_medium >= 0) shape->addChild("exterior", rs->media[s.exterior_medium]);
        shape->configure();
        rs->shapeIndex[shape.get()] = i;
        rs->shapes.push_back(shape);
        rs->scene->addChild(shape);
    }
    return rs.release();
}

static void setIntegrator(RefScene *rs, const B200pgIntegratorParams *P, const char *plugin = nullptr, int deviceCount = 1) {
    if (!plugin && rs->haveInteg && !rs->pluginInteg && memcmp(&rs->lastParams, P, sizeof(*P)) == 0) return;
    Properties p(plugin ? plugin : (P->volumetric ? "progressivevolpath" : "progressivepath"));
    if (plugin) {  // the reference-side binding of libb200pg.so (integration/b200guidedpath.cpp): its own parameters
        p.setBoolean("guiding", P->guiding != 0);
        p.setInteger("trainingProgressions", P->training_progressions);
        p.setFloat("guidingProbability", P->guiding_probability);
        p.setInteger("maxComponents", P->guide_max_components);
        p.setInteger("maxSamplesPerCell", P->guide_max_cell_samples);
        p.setBoolean("discardTrainingSamples", P->guide_train_discard_film != 0);
        p.setBoolean("guidedDistance", P->guided_distance != 0);
        p.setBoolean("volumetric", P->volumetric != 0);
        p.setInteger("deviceCount", deviceCount);
    }
    p.setInteger("maxDepth", P->max_depth);
    p.setInteger("rrDepth", P->rr_depth);
    p.setBoolean("strictNormals", P->strict_normals != 0);
    p.setBoolean("hideEmitters", P->hide_emitters != 0);
    p.setInteger("samplesPerProgression", P->samples_per_progression > 0 ? P->samples_per_progression : 1);
    p.setInteger("maxRenderTime", P->max_render_time);
    if (std::isfinite(P->max_component_value)) p.setFloat("maxComponentValue", P->max_component_value);
    p.setBoolean("useNee", P->use_nee != 0);  // progressive_path.cpp:117, progressive_volpath.cpp:82
    rs->pluginInteg = plugin != nullptr;
    // progressive_path.cpp:340 registers the class under MonteCarloIntegrator, so that is the type to ask the plugin manager for
    ref<SamplingIntegrator> integ = static_cast<SamplingIntegrator *>(create<MonteCarloIntegrator>(p).get());
    integ->configure();
    rs->integ = integ;
    rs->scene->setIntegrator(integ);
    if (!rs->haveInteg) {
        if (getenv("REF_HARNESS_VERBOSE")) fprintf(stderr, "configure\n");
        rs->scene->configure();
        if (getenv("REF_HARNESS_VERBOSE")) fprintf(stderr, "initialize\n");
        rs->scene->initialize();
        if (getenv("REF_HARNESS_VERBOSE")) fprintf(stderr, "initialized\n");  // builds the ShapeKDTree (scene.cpp:260-330)
    }
    rs->lastParams = *P;
    rs->haveInteg = true;
}

static void ensureBuilt(RefScene *rs) {
    if (rs->haveInteg) return;
    B200pgIntegratorParams P;
    memset(&P, 0, sizeof(P));
    P.max_depth = -1;
    P.rr_depth = 5;
    P.samples_per_progression = 1;
    P.max_component_value = std::numeric_limits<float>::infinity();
    P.use_nee = 1;
    setIntegrator(rs, &P);
}

#define REF_TRY try {
#define REF_CATCH(ret)                                  \
    }                                                   \
    catch (const std::exception &e) {                   \
        g_err = e.what();                               \
        fprintf(stderr, "ref_harness: %s\n", e.what()); \
        return ret;                                     \
    }

static uint32_t globalPrim(const RefScene *rs, const Intersection &its) {
    auto it = rs->shapeIndex.find(its.shape);
    if (it == rs->shapeIndex.end()) return 0xFFFFFFFEu;
    const bool mesh = its.shape->getClass()->derivesFrom(MTS_CLASS(TriMesh));
    return rs->primOffset[it->second] + (mesh ? its.primIndex : 0u);
}

static int renderCore(RefScene *rs, int first_sample, int n_samples, float *film, int nthreads, int independent, double *seconds,
                      int *spp_done, int repeat) {
    ensureWorkers(nthreads);
    Scheduler *sched = Scheduler::getInstance();
    ref<Sampler> sampler;
    if (independent) {
        Properties ps("independent");
        ps.setInteger("sampleCount", n_samples);
        sampler = create<Sampler>(ps);
        sampler->configure();
    } else {
        sampler = new ReplaySampler(rs->seed, rs->width, (size_t)n_samples, (size_t)first_sample);
    }
    rs->scene->setSampler(sampler);
    ref<RenderQueue> queue = new RenderQueue();
    // the registrations RenderJob's constructor would make (renderjob.cpp:38-72), made here so that the ids are known
    const int sceneID = sched->registerResource(rs->scene), sensorID = sched->registerResource(rs->sensor);
    std::vector<SerializableObject *> samplers(sched->getCoreCount());
    for (size_t i = 0; i < samplers.size(); ++i) {
        ref<Sampler> c = sampler->clone();
        c->incRef();
        samplers[i] = c.get();
    }
    const int samplerID = sched->registerMultiResource(samplers);
    for (size_t i = 0; i < samplers.size(); ++i) samplers[i]->decRef();
    ref<RenderJob> job = new RenderJob("ref", rs->scene, queue, sceneID, sensorID, samplerID, false);
    // RenderJob::run (renderjob.cpp:84-112) without the develop step
    if (!rs->scene->preprocess(queue, job, sceneID, sensorID, samplerID)) throw std::runtime_error("preprocess failed");
    // `repeat` > 1 (timing runs): Scene::render again on the same per-pixel samplers (preprocess allocates one sampler per pixel,
    // progressiveintegrator.cpp:43-51 -- seconds for a 1024^2 film -- and is not part of the timed region)
    // repeat < 0: one untimed warm-up render, then renders until -repeat milliseconds have passed (at least two)
    int done = 0;
    if (repeat < 0 && !rs->scene->render(queue, job, sceneID, sensorID, samplerID)) throw std::runtime_error("render failed");
    auto t0 = std::chrono::steady_clock::now();
    double elapsed = 0;
    while (true) {
        if (!rs->scene->render(queue, job, sceneID, sensorID, samplerID)) throw std::runtime_error("render failed");
        ++done;
        elapsed = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        if (repeat >= 0 ? done >= std::max(repeat, 1) : (done >= 2 && elapsed * 1e3 >= -repeat)) break;
    }
    if (seconds) *seconds = elapsed;
    if (spp_done) *spp_done = n_samples * done;
    queue->removeJob(job, false);
    sched->unregisterResource(sceneID);
    sched->unregisterResource(sensorID);
    sched->unregisterResource(samplerID);
    if (film) {
        ImageBlock *st = rs->film->getStorage();
        if (!st) throw std::runtime_error("film has no storage");
        const int b = st->getBorderSize(), sx = rs->width + 2 * b;
        const Float *data = st->getBitmap()->getFloatData();
        for (int y = 0; y < rs->height; ++y)
            for (int x = 0; x < rs->width; ++x)
                for (int k = 0; k < 5; ++k) film[((size_t)y * rs->width + x) * 5 + k] = data[((size_t)(y + b) * sx + (x + b)) * 5 + k];
    }
    return 0;
}


// "name|kind|value;..." -> Properties, the way SceneHandler::endElement fills them from XML attributes (scenehandler.cpp:296-470).
// kinds: f float, i integer, b boolean, s string, c "r,g,b" colour, x transform = '/'-separated ops applied in order with the
// handler's rule m_transform = op * m_transform (:348-440) and the reference's own Transform factories:
//   t:x,y,z   r:x,y,z,angle(degrees)   s:x,y,z   l:ox,oy,oz,tx,ty,tz[,ux,uy,uz] (lookat; no up -> coordinateSystem, :391-396)   m:16 values
static Properties parseProps(const char *type, const char *props) {
    Properties p(type);
    for (const std::string &item : tokenize(props ? props : "", ";")) {
        std::vector<std::string> f = tokenize(item, "|");
        if (f.size() != 3) throw std::runtime_error("bad property \"" + item + "\"");
        const std::string &name = f[0], &kind = f[1], &value = f[2];
        if (kind == "f") p.setFloat(name, (Float)atof(value.c_str()));
        else if (kind == "i") p.setInteger(name, atoi(value.c_str()));
        else if (kind == "b") p.setBoolean(name, value == "true");
        else if (kind == "s") p.setString(name, value);
        else if (kind == "c") {
            float c[3];
            if (sscanf(value.c_str(), "%f,%f,%f", &c[0], &c[1], &c[2]) != 3) throw std::runtime_error("bad colour");
            p.setSpectrum(name, rgb(c));
        } else if (kind == "x") {
            Transform trafo;
            for (const std::string &op : tokenize(value, "/")) {
                std::vector<Float> v;
                for (const std::string &t : tokenize(op.substr(2), ",")) v.push_back((Float)atof(t.c_str()));
                switch (op[0]) {
                    case 't': trafo = Transform::translate(Vector(v.at(0), v.at(1), v.at(2))) * trafo; break;
                    case 'r': trafo = Transform::rotate(Vector(v.at(0), v.at(1), v.at(2)), v.at(3)) * trafo; break;
                    case 's': trafo = Transform::scale(Vector(v.at(0), v.at(1), v.at(2))) * trafo; break;
                    case 'l': {
                        Point o(v.at(0), v.at(1), v.at(2)), t(v.at(3), v.at(4), v.at(5));
                        Vector up(0.0f);
                        if (v.size() >= 9) up = Vector(v[6], v[7], v[8]);
                        if (up.lengthSquared() == 0) {
                            Vector unused;
                            coordinateSystem(normalize(t - o), up, unused);
                        }
                        trafo = Transform::lookAt(o, t, up) * trafo;
                        break;
                    }
                    case 'm': {
                        Matrix4x4 m;
                        for (int i = 0; i < 16; ++i) m.m[i / 4][i % 4] = v.at(i);
                        trafo = Transform(m) * trafo;
                        break;
                    }
                    default: throw std::runtime_error("bad transform op");
                }
            }
            p.setTransform(name, trafo);
        } else throw std::runtime_error("bad property kind");
    }
    return p;
}


extern "C" {

const char *ref_last_error() { return g_err.c_str(); }

void *ref_scene_create(const B200pgSceneDesc *desc) {
    REF_TRY
    return buildScene(desc);
    REF_CATCH(nullptr)
}
void ref_scene_destroy(void *s) { delete (RefScene *)s; }

// The shapes of the description WITHOUT its sensor: Scene::configure then adds its fallback camera (scene.cpp:272-305: 45 degree
// perspective on the -z side of the shapes' bounding box, clip planes from its extents) with the default film. The handle's
// sensor / film are that camera's afterwards; wh = film size.
void *ref_scene_create_without_sensor(const B200pgSceneDesc *desc, int *wh) {
    REF_TRY
    g_noSensor = true;
    RefScene *rs = nullptr;
    try {
        rs = buildScene(desc);
    } catch (...) {
        g_noSensor = false;
        throw;
    }
    g_noSensor = false;
    ensureBuilt(rs);
    rs->sensor = rs->scene->getSensor();
    rs->film = rs->sensor->getFilm();
    rs->width = rs->film->getSize().x;
    rs->height = rs->film->getSize().y;
    wh[0] = rs->width;
    wh[1] = rs->height;
    return rs;
    REF_CATCH(nullptr)
}

// ShapeKDTree statistics of the reference's own SAH build: out = {node count is not exported by the class; shapes, primitives}
int ref_kd_info(void *s, uint64_t *out) {
    REF_TRY
    RefScene *rs = (RefScene *)s;
    ensureBuilt(rs);
    out[0] = rs->scene->getKDTree()->getShapes().size();
    out[1] = rs->scene->getKDTree()->getPrimitiveCount();
    return 0;
    REF_CATCH(-1)
}

// Scene::rayIntersect (closest hit, skdtree.cpp:112-142 -> sahkdtree3.h rayIntersectHavran) or the shadow-ray overload.
// rays n*8 (o, mint, d, maxt). out n*18 = {t, p.xyz, uv.xy, geoFrame.n, shFrame.n, shFrame.s, dpdu}, t = inf for a miss;
// prim n = global primitive id (shape prefix + triangle index), 0xFFFFFFFF for a miss. shadow: prim = 0 / 0xFFFFFFFF only.
int ref_intersect(void *s, const float *rays, size_t n, int shadow, float *out, uint32_t *prim, int nthreads) {
    REF_TRY
    RefScene *rs = (RefScene *)s;
    ensureBuilt(rs);
    const Scene *scene = rs->scene;
    if (nthreads <= 0) nthreads = omp_get_max_threads();
#pragma omp parallel for num_threads(nthreads) schedule(dynamic, 4096)
    for (long long i = 0; i < (long long)n; ++i) {
        const float *r = rays + 8 * i;
        Ray ray(Point(r[0], r[1], r[2]), Vector(r[4], r[5], r[6]), r[3], r[7], 0.0f);
        if (shadow) {
            prim[i] = scene->rayIntersect(ray) ? 0u : 0xFFFFFFFFu;
            continue;
        }
        Intersection its;
        float *o = out ? out + 18 * i : nullptr;
        if (o)
            for (int k = 0; k < 18; ++k) o[k] = 0.0f;
        if (!scene->rayIntersect(ray, its)) {
            if (o) o[0] = std::numeric_limits<float>::infinity();
            prim[i] = 0xFFFFFFFFu;
            continue;
        }
        prim[i] = globalPrim(rs, its);
        if (!o) continue;
        o[0] = its.t;
        o[1] = its.p.x; o[2] = its.p.y; o[3] = its.p.z;
        o[4] = its.uv.x; o[5] = its.uv.y;
        o[6] = its.geoFrame.n.x; o[7] = its.geoFrame.n.y; o[8] = its.geoFrame.n.z;
        o[9] = its.shFrame.n.x; o[10] = its.shFrame.n.y; o[11] = its.shFrame.n.z;
        o[12] = its.shFrame.s.x; o[13] = its.shFrame.s.y; o[14] = its.shFrame.s.z;
        o[15] = its.dpdu.x; o[16] = its.dpdu.y; o[17] = its.dpdu.z;
    }
    return 0;
    REF_CATCH(-1)
}

// PerspectiveCamera::sampleRay for film positions pos n*2 (pixel units): rays n*8
int ref_camera_rays(void *s, const float *pos, size_t n, float *rays) {
    REF_TRY
    RefScene *rs = (RefScene *)s;
    for (size_t i = 0; i < n; ++i) {
        Ray r;
        rs->sensor->sampleRay(r, Point2(pos[2 * i], pos[2 * i + 1]), Point2(0.5f), 0.5f);
        float *o = rays + 8 * i;
        o[0] = r.o.x; o[1] = r.o.y; o[2] = r.o.z; o[3] = r.mint;
        o[4] = r.d.x; o[5] = r.d.y; o[6] = r.d.z; o[7] = r.maxt;
    }
    return 0;
    REF_CATCH(-1)
}

// BSDF::eval / pdf for (wi, wo) pairs and BSDF::sample for (wi, u), all in the local frame (the calling convention of
// src/tests/test_chisquare.cpp's BSDFAdapter: an Intersection with identity frames).
int ref_bsdf(void *s, int bsdfIndex, const float *wi, const float *wo, const float *u, size_t n, float *out_eval, float *out_pdf,
             float *out_wo, float *out_weight, float *out_spdf, uint32_t *out_flags) {
    REF_TRY
    RefScene *rs = (RefScene *)s;
    if (bsdfIndex < 0 || bsdfIndex >= (int)rs->bsdfs.size()) return -1;
    const BSDF *bsdf = rs->bsdfs[bsdfIndex];
    for (size_t i = 0; i < n; ++i) {
        Intersection its;
        its.p = Point(0.0f);
        its.geoFrame = Frame(Normal(0, 0, 1));
        its.shFrame = its.geoFrame;
        its.uv = Point2(0.5f);
        its.hasUVPartials = false;
        its.time = 0;
        Vector vi(wi[3 * i], wi[3 * i + 1], wi[3 * i + 2]), vo(wo[3 * i], wo[3 * i + 1], wo[3 * i + 2]);
        its.wi = vi;
        {
            BSDFSamplingRecord bRec(its, vi, vo, ERadiance);
            // discrete (delta) components are not part of eval / pdf with the solid-angle measure
            Spectrum e = bsdf->eval(bRec, ESolidAngle);
            out_eval[3 * i] = e[0]; out_eval[3 * i + 1] = e[1]; out_eval[3 * i + 2] = e[2];
            out_pdf[i] = bsdf->pdf(bRec, ESolidAngle);
        }
        {
            BSDFSamplingRecord bRec(its, NULL, ERadiance);
            bRec.wi = vi;
            Float pdf = 0;
            Spectrum w = bsdf->sample(bRec, pdf, Point2(u[2 * i], u[2 * i + 1]));
            if (w.isZero()) {
                bRec.wo = Vector(0.0f);
                pdf = 0;
                bRec.sampledType = 0;
            }
            out_wo[3 * i] = bRec.wo.x; out_wo[3 * i + 1] = bRec.wo.y; out_wo[3 * i + 2] = bRec.wo.z;
            out_weight[3 * i] = w[0]; out_weight[3 * i + 1] = w[1]; out_weight[3 * i + 2] = w[2];
            out_spdf[i] = pdf;
            out_flags[i] = bRec.sampledType;
        }
    }
    return 0;
    REF_CATCH(-1)
}

// Scene::sampleEmitterDirect(dRec, u, testVisibility = false) for u n*2, or Scene::pdfEmitterDirect for given directions
// (the first surface along (ref, d) must be an emitter; the record is filled from the hit as progressive_path.cpp:243-262)
int ref_emitter_direct(void *s, const float *refp, const float *refN, const float *u, float *d, size_t n, float *out_dist,
                       float *out_pdf, float *out_value) {
    REF_TRY
    RefScene *rs = (RefScene *)s;
    ensureBuilt(rs);
    const Scene *scene = rs->scene;
    for (size_t i = 0; i < n; ++i) {
        DirectSamplingRecord dRec(Point(refp[0], refp[1], refp[2]), 0.0f);
        dRec.refN = Normal(refN[0], refN[1], refN[2]);
        if (u) {
            Spectrum v = scene->sampleEmitterDirect(dRec, Point2(u[2 * i], u[2 * i + 1]), false);
            d[3 * i] = dRec.d.x; d[3 * i + 1] = dRec.d.y; d[3 * i + 2] = dRec.d.z;
            if (out_dist) out_dist[i] = dRec.dist;
            if (out_pdf) out_pdf[i] = dRec.pdf;
            if (out_value) { out_value[3 * i] = v[0]; out_value[3 * i + 1] = v[1]; out_value[3 * i + 2] = v[2]; }
        } else {
            Ray ray(dRec.ref, Vector(d[3 * i], d[3 * i + 1], d[3 * i + 2]), 0.0f);
            Intersection its;
            Float pdf = 0;
            if (scene->rayIntersect(ray, its) && its.isEmitter()) {
                dRec.setQuery(ray, its);
                pdf = scene->pdfEmitterDirect(dRec);
                if (out_dist) out_dist[i] = its.t;
            } else if (out_dist) {
                out_dist[i] = std::numeric_limits<float>::infinity();
            }
            out_pdf[i] = pdf;
        }
    }
    return 0;
    REF_CATCH(-1)
}

// The integrator's Li for the camera samples (pixel[i], sample[i]): the loop body of
// ProgressiveMonteCarloIntegrator::renderBlock (progressiveintegrator.cpp:253-283) up to and without the clamp / film splat.
int ref_radiance(void *s, const B200pgIntegratorParams *P, const uint32_t *pixel, const uint32_t *sample, size_t n, float *out_rgb,
                 float *out_pos) {
    REF_TRY
    RefScene *rs = (RefScene *)s;
    setIntegrator(rs, P);
    const Scene *scene = rs->scene;
    const Sensor *sensor = rs->sensor;
    const SamplingIntegrator *integ = rs->integ;
    const int W = rs->width;
    const Float diffScale = 1.0f / std::sqrt((Float)(P->samples_per_progression > 0 ? P->samples_per_progression : 1));
#pragma omp parallel
    {
        ref<ReplaySampler> sampler = new ReplaySampler(rs->seed, W, rs->sampleCount, 0);
#pragma omp for schedule(dynamic, 256)
        for (long long i = 0; i < (long long)n; ++i) {
            sampler->setStream(pixel[i], sample[i]);
            RadianceQueryRecord rRec(scene, sampler);
            uint32_t queryType = RadianceQueryRecord::ESensorRay;
            if (!sensor->getFilm()->hasAlpha()) queryType &= ~RadianceQueryRecord::EOpacity;
            rRec.newQuery(queryType, sensor->getMedium());
            rRec.pixelId = pixel[i];
            Point2i offset(pixel[i] % W, pixel[i] / W);
            Point2 samplePos(Point2(offset) + Vector2(rRec.nextSample2D()));
            RayDifferential sensorRay;
            Spectrum spec = sensor->sampleRayDifferential(sensorRay, samplePos, Point2(0.5f), 0.5f);
            sensorRay.scaleDifferential(diffScale);
            spec *= integ->Li(sensorRay, rRec);
            out_rgb[3 * i] = spec[0]; out_rgb[3 * i + 1] = spec[1]; out_rgb[3 * i + 2] = spec[2];
            if (out_pos) { out_pos[2 * i] = samplePos.x; out_pos[2 * i + 1] = samplePos.y; }
        }
    }
    return 0;
    REF_CATCH(-1)
}

// ImageBlock::put (imageblock.h:131-197) with the film's reconstruction filter on a film-sized block with the film's border;
// film: H*W*5 accumulators (+=), the layout of orc_film_splat.
int ref_film_splat(void *s, const float *pos, const float *rgbv, size_t n, float *film) {
    REF_TRY
    RefScene *rs = (RefScene *)s;
    const ReconstructionFilter *rf = rs->film->getReconstructionFilter();
    ref<ImageBlock> blk = new ImageBlock(Bitmap::ESpectrumAlphaWeight, Vector2i(rs->width, rs->height), rf);
    blk->setOffset(Point2i(0, 0));
    blk->clear();
    for (size_t i = 0; i < n; ++i) {
        Spectrum v;
        v.fromLinearRGB(rgbv[3 * i], rgbv[3 * i + 1], rgbv[3 * i + 2]);
        blk->put(Point2(pos[2 * i], pos[2 * i + 1]), v, 1.0f);
    }
    const int b = blk->getBorderSize(), sx = rs->width + 2 * b;
    const Float *data = blk->getBitmap()->getFloatData();
    for (int y = 0; y < rs->height; ++y)
        for (int x = 0; x < rs->width; ++x)
            for (int k = 0; k < 5; ++k) film[((size_t)y * rs->width + x) * 5 + k] += data[((size_t)(y + b) * sx + (x + b)) * 5 + k];
    return 0;
    REF_CATCH(-1)
}

// The reference's render loop itself: Scene::preprocess + Scene::render (= ProgressiveMonteCarloIntegrator::render ->
// renderSamples / renderTime -> BlockedRenderProcess on `nthreads` LocalWorkers -> renderBlock -> Film::put), everything but
// Film::develop. Renders samples [first_sample, first_sample + n_samples) of every pixel with the replay sampler
// (independent = 0) or the reference's own `independent` sampler (independent = 1: the stock configuration; timing runs).
// film: H*W*5 (the HDRFilm storage without its border, overwritten). seconds: wall time of the Scene::render calls alone.
int ref_render(void *s, const B200pgIntegratorParams *P, int first_sample, int n_samples, float *film, int nthreads, int independent,
               double *seconds, int *spp_done, int repeat) {
    REF_TRY
    RefScene *rs = (RefScene *)s;
    setIntegrator(rs, P);
    return renderCore(rs, first_sample, n_samples, film, nthreads, independent, seconds, spp_done, repeat);
    REF_CATCH(-1)
}

// Scene::preprocess + Scene::render + Integrator::postprocess with the integrator the reference's PluginManager creates from
// plugins/<plugin>.so -- integration/b200guidedpath.cpp, the reference-side binding of libb200pg.so -- on a scene whose source
// file is `xml_path` (the binding hands that file to the GPU library). film: the reference's HDRFilm storage afterwards.
int ref_render_plugin(void *s, const B200pgIntegratorParams *P, const char *plugin, const char *xml_path, int device_count,
                      float *film, double *seconds) {
    REF_TRY
    RefScene *rs = (RefScene *)s;
    setIntegrator(rs, P, plugin, device_count);
    rs->scene->setSourceFile(fs::path(xml_path ? xml_path : ""));
    int spp = 0;
    const int rc = renderCore(rs, 0, rs->sampleCount, film, 1, 0, seconds, &spp, 1);
    rs->scene->getIntegrator()->postprocess(rs->scene, nullptr, nullptr, -1, -1, -1);
    return rc;
    REF_CATCH(-1)
}
// The film hand-off of the reference-side plugin on its own (integration/b200guidedpath.cpp: b200guidedpath_put_film), so that it
// can be checked without a device: clears the scene's film, lets the plugin put `rgbaw` (H*W*5) into it, reads the storage back.
int ref_plugin_put_film(void *s, const float *rgbaw, float *film) {
    REF_TRY
    RefScene *rs = (RefScene *)s;
    const FileResolver *resolver = Thread::getThread()->getFileResolver();
    fs::path path = resolver->resolve(fs::path("plugins") / "b200guidedpath.so");
    void *h = dlopen(path.string().c_str(), RTLD_LAZY | RTLD_LOCAL);
    if (!h) throw std::runtime_error(std::string("dlopen: ") + dlerror());
    typedef void (*PutFilm)(Film *, const float *);
    PutFilm put = (PutFilm)dlsym(h, "b200guidedpath_put_film");
    if (!put) throw std::runtime_error("b200guidedpath_put_film not exported");
    rs->film->clear();
    put(rs->film, rgbaw);
    ImageBlock *st = rs->film->getStorage();
    const int b = st->getBorderSize(), sx = rs->width + 2 * b;
    const Float *data = st->getBitmap()->getFloatData();
    for (int y = 0; y < rs->height; ++y)
        for (int x = 0; x < rs->width; ++x)
            for (int k = 0; k < 5; ++k) film[((size_t)y * rs->width + x) * 5 + k] = data[((size_t)(y + b) * sx + (x + b)) * 5 + k];
    return 0;
    REF_CATCH(-1)
}

// GridDataSource::lookupFloat (gridvolume.cpp:337-388)
int ref_grid_lookup(void *s, int medium, const float *p, size_t n, float *out) {
    REF_TRY
    RefScene *rs = (RefScene *)s;
    if (medium < 0 || medium >= (int)rs->densities.size()) return -1;
    for (size_t i = 0; i < n; ++i) out[i] = rs->densities[medium]->lookupFloat(Point(p[3 * i], p[3 * i + 1], p[3 * i + 2]));
    return 0;
    REF_CATCH(-1)
}

// HeterogeneousMedium::sampleDistance, then evalTransmittance (heterogeneous.cpp), both drawing from the replay stream
// (seed, pixel = i, sample 0) one after the other -- the order orc_medium_sample uses.
int ref_medium_sample(void *s, int medium, const float *rays, size_t n, float *out_t, float *out_success, float *out_tr) {
    REF_TRY
    RefScene *rs = (RefScene *)s;
    if (medium < 0 || medium >= (int)rs->media.size()) return -1;
    const Medium *M = rs->media[medium];
    ref<ReplaySampler> sampler = new ReplaySampler(rs->seed, rs->width, 2, 0);
    for (size_t i = 0; i < n; ++i) {
        const float *r = rays + 8 * i;
        Ray ray(Point(r[0], r[1], r[2]), Vector(r[4], r[5], r[6]), r[3], r[7], 0.0f);
        MediumSamplingRecord mRec;
        sampler->setStream((uint32_t)i, 0);
        bool ok = M->sampleDistance(ray, mRec, sampler);
        out_t[i] = ok ? mRec.t : std::numeric_limits<float>::infinity();
        if (out_success) out_success[i] = ok ? mRec.pdfSuccess : mRec.pdfFailure;
        Spectrum tr = M->evalTransmittance(ray, sampler);
        out_tr[i] = tr[0];
    }
    return 0;
    REF_CATCH(-1)
}

// PhaseFunction::eval for (wi, wo) and PhaseFunction::sample for (wi, u) (hg.cpp, isotropic.cpp)
int ref_phase(void *s, int medium, const float *wi, const float *wo, const float *u, size_t n, float *out_eval, float *out_wo,
              float *out_pdf) {
    REF_TRY
    RefScene *rs = (RefScene *)s;
    if (medium < 0 || medium >= (int)rs->media.size()) return -1;
    const PhaseFunction *ph = rs->media[medium]->getPhaseFunction();
    struct FixedSampler : public Sampler {
        Point2 v;
        FixedSampler() : Sampler(Properties()) {}
        Float next1D() { return v.x; }
        Point2 next2D() { return v; }
    };
    ref<FixedSampler> fs = new FixedSampler();
    for (size_t i = 0; i < n; ++i) {
        MediumSamplingRecord mRec;
        Vector vi(wi[3 * i], wi[3 * i + 1], wi[3 * i + 2]), vo(wo[3 * i], wo[3 * i + 1], wo[3 * i + 2]);
        PhaseFunctionSamplingRecord pRec(mRec, vi, vo);
        out_eval[i] = ph->eval(pRec);
        PhaseFunctionSamplingRecord sRec(mRec, vi);
        fs->v = Point2(u[2 * i], u[2 * i + 1]);
        Float pdf = 0;
        ph->sample(sRec, pdf, fs);
        out_wo[3 * i] = sRec.wo.x; out_wo[3 * i + 1] = sRec.wo.y; out_wo[3 * i + 2] = sRec.wo.z;
        out_pdf[i] = pdf;
    }
    return 0;
    REF_CATCH(-1)
}

// ---------------------------------------------------------------------------------------------------------------------------
// Mesh ingestion: the reference's own loaders -- the `obj` plugin (src/shapes/obj.cpp) and
// TriMesh(Stream *, int) for `.serialized` files (trimesh.cpp:79-270; the `serialized` plugin adds only an offset cache on top
// of it) -- followed by TriMesh::configure (normals: trimesh.cpp:608-681). Returned as plain arrays.
// ---------------------------------------------------------------------------------------------------------------------------
struct RefMeshes {
    ref<Shape> owner;
    std::vector<ref<TriMesh>> meshes;
};

void *ref_mesh_load(const char *kind, const char *path, int shape_index, int face_normals, int flip_normals, const float *to_world) {
    REF_TRY
    ensureInit();
    std::unique_ptr<RefMeshes> rm(new RefMeshes());
    if (std::string(kind) == "obj") {
        Properties p("obj");
        p.setString("filename", path);
        p.setBoolean("faceNormals", face_normals != 0);
        p.setBoolean("flipNormals", flip_normals != 0);
        if (to_world) p.setTransform("toWorld", xform(to_world));
        ref<Shape> shape = create<Shape>(p);
        shape->configure();
        rm->owner = shape;
        for (int i = 0;; ++i) {
            Shape *e = shape->getElement(i);
            if (!e) break;
            TriMesh *m = static_cast<TriMesh *>(e);
            m->configure();
            rm->meshes.push_back(m);
        }
    } else {
        ref<FileStream> fs = new FileStream(path, FileStream::EReadOnly);
        // the serialized plugin seeks to the shape's offset from the dictionary at the end of the file (serialized.cpp) before
        // handing the stream to TriMesh; with idx > 0 TriMesh::TriMesh(Stream *, int) does the same through readOffset
        ref<TriMesh> m = new TriMesh(fs, shape_index);
        m->configure();
        rm->meshes.push_back(m);
    }
    return rm.release();
    REF_CATCH(nullptr)
}
void ref_mesh_destroy(void *h) { delete (RefMeshes *)h; }
int ref_mesh_count(void *h) { return (int)((RefMeshes *)h)->meshes.size(); }
int ref_mesh_info(void *h, int i, uint64_t *out /* vertices, triangles, has normals, has texcoords */) {
    const TriMesh *m = ((RefMeshes *)h)->meshes[i];
    out[0] = m->getVertexCount();
    out[1] = m->getTriangleCount();
    out[2] = m->getVertexNormals() != NULL;
    out[3] = m->getVertexTexcoords() != NULL;
    return 0;
}
int ref_mesh_get(void *h, int i, float *pos, float *nrm, float *uv, uint32_t *idx) {
    const TriMesh *m = ((RefMeshes *)h)->meshes[i];
    for (size_t v = 0; v < m->getVertexCount(); ++v) {
        const Point &p = m->getVertexPositions()[v];
        pos[3 * v] = p.x; pos[3 * v + 1] = p.y; pos[3 * v + 2] = p.z;
        if (nrm && m->getVertexNormals()) {
            const Normal &n = m->getVertexNormals()[v];
            nrm[3 * v] = n.x; nrm[3 * v + 1] = n.y; nrm[3 * v + 2] = n.z;
        }
        if (uv && m->getVertexTexcoords()) {
            const Point2 &t = m->getVertexTexcoords()[v];
            uv[2 * v] = t.x; uv[2 * v + 1] = t.y;
        }
    }
    for (size_t t = 0; t < m->getTriangleCount(); ++t)
        for (int k = 0; k < 3; ++k) idx[3 * t + k] = m->getTriangles()[t].idx[k];
    return 0;
}

// A BSDF created the way the scene loader would from XML -- plugin name + named properties, nothing filled in by this file --
// so that the product's XML reader can be held to the reference's own defaults and property semantics (named IORs, alpha vs
// alphaU / alphaV, distribution names, ...). props: "name|kind|value;..." with kind f (float), i (integer), b (boolean:
// true / false), s (string), c (r,g,b colour). twosided != 0 wraps the result in the `twosided` plugin. Evaluated like ref_bsdf.
int ref_bsdf_from_props(const char *type, const char *props, int twosided, const float *wi, const float *wo, const float *u, size_t n,
                        float *out_eval, float *out_pdf, float *out_wo, float *out_weight, float *out_spdf, uint32_t *out_flags) {
    REF_TRY
    ensureInit();
    Properties p = parseProps(type, props);
    ref<BSDF> bsdf = create<BSDF>(p);
    bsdf->configure();
    if (twosided) {
        ref<BSDF> outer = create<BSDF>(Properties("twosided"));
        outer->addChild(bsdf);
        outer->configure();
        bsdf = outer;
    }
    RefScene tmp;
    tmp.bsdfs.push_back(bsdf);
    return ref_bsdf(&tmp, 0, wi, wo, u, n, out_eval, out_pdf, out_wo, out_weight, out_spdf, out_flags);
    REF_CATCH(-1)
}

// A `perspective` sensor from named properties (+ an hdrfilm of the given size): rays for film positions pos (pixel units)
int ref_sensor_rays_from_props(const char *props, int width, int height, const float *pos, size_t n, float *rays) {
    REF_TRY
    ensureInit();
    Properties pf("hdrfilm");
    pf.setInteger("width", width);
    pf.setInteger("height", height);
    pf.setBoolean("banner", false);
    ref<Film> film = create<Film>(pf);
    film->configure();
    ref<Sensor> sensor = create<Sensor>(parseProps("perspective", props));
    sensor->addChild(film);
    sensor->configure();
    for (size_t i = 0; i < n; ++i) {
        Ray r;
        sensor->sampleRay(r, Point2(pos[2 * i], pos[2 * i + 1]), Point2(0.5f), 0.5f);
        float *o = rays + 8 * i;
        o[0] = r.o.x; o[1] = r.o.y; o[2] = r.o.z; o[3] = r.mint;
        o[4] = r.d.x; o[5] = r.d.y; o[6] = r.d.z; o[7] = r.maxt;
    }
    return 0;
    REF_CATCH(-1)
}

// One shape plugin (rectangle, cube, obj) from named properties in a ShapeKDTree of its own: hit records as ref_intersect
int ref_shape_from_props(const char *plugin, const char *props, const float *rays, size_t n, float *out) {
    REF_TRY
    ensureInit();
    ref<Shape> shape = create<Shape>(parseProps(plugin, props));
    ref<BSDF> bsdf = create<BSDF>(Properties("diffuse"));
    bsdf->configure();
    shape->addChild(bsdf);
    shape->configure();
    ref<ShapeKDTree> tree = new ShapeKDTree();
    if (shape->getClass()->getName() == "WavefrontOBJ") {
        for (int i = 0;; ++i) {
            Shape *e = shape->getElement(i);
            if (!e) break;
            e->configure();
            tree->addShape(e);
        }
    } else {
        tree->addShape(shape);
    }
    tree->build();
    for (size_t i = 0; i < n; ++i) {
        const float *r = rays + 8 * i;
        Ray ray(Point(r[0], r[1], r[2]), Vector(r[4], r[5], r[6]), r[3], r[7], 0.0f);
        Intersection its;
        float *o = out + 18 * i;
        for (int k = 0; k < 18; ++k) o[k] = 0.0f;
        if (!tree->rayIntersect(ray, its)) {
            o[0] = std::numeric_limits<float>::infinity();
            continue;
        }
        o[0] = its.t;
        o[1] = its.p.x; o[2] = its.p.y; o[3] = its.p.z;
        o[4] = its.uv.x; o[5] = its.uv.y;
        o[6] = its.geoFrame.n.x; o[7] = its.geoFrame.n.y; o[8] = its.geoFrame.n.z;
        o[9] = its.shFrame.n.x; o[10] = its.shFrame.n.y; o[11] = its.shFrame.n.z;
        o[12] = its.shFrame.s.x; o[13] = its.shFrame.s.y; o[14] = its.shFrame.s.z;
        o[15] = its.dpdu.x; o[16] = its.dpdu.y; o[17] = its.dpdu.z;
    }
    return 0;
    REF_CATCH(-1)
}

// Defaults of the reference's objects created without any property: out = {film width, film height, reconstruction filter
// radius, sampler sampleCount, area light samplingWeight}
int ref_defaults(float *out) {
    REF_TRY
    ensureInit();
    ref<Film> film = create<Film>(Properties("hdrfilm"));
    film->configure();
    ref<Sampler> sampler = create<Sampler>(Properties("independent"));
    sampler->configure();
    Properties pe("area");
    pe.setSpectrum("radiance", Spectrum(1.0f));
    ref<Emitter> em = create<Emitter>(pe);
    out[0] = (float)film->getSize().x;
    out[1] = (float)film->getSize().y;
    out[2] = film->getReconstructionFilter()->getRadius();
    out[3] = (float)sampler->getSampleCount();
    out[4] = em->getSamplingWeight();
    return 0;
    REF_CATCH(-1)
}

// The reference's COUNTING traversal (rayIntersectHavranCollectStatistics, sahkdtree3.h:330-429: the variant it uses to report
// kd-tree cost, without mailboxing) on the scene's kd-tree, with the ray interval set up as ShapeKDTree::rayIntersect does
// (skdtree.cpp:112-142). The method is protected in SAHKDTree3D, hence the accessor subclass. Sums over the n rays:
// counters = {inner nodes traversed (numTraversals), leaf index entries visited (numIntersections), rays that hit}.
struct KDProbe : public ShapeKDTree {
    void count(const Ray &ray, uint64_t *c) const {
        uint8_t temp[MTS_KD_INTERSECTION_TEMP];
        Float mint, maxt, t = std::numeric_limits<Float>::infinity();
        if (!m_aabb.rayIntersect(ray, mint, maxt)) return;
        Float rayMinT = ray.mint;
        if (rayMinT == Epsilon) rayMinT *= std::max(std::max(std::max(std::abs(ray.o.x), std::abs(ray.o.y)), std::abs(ray.o.z)), Epsilon);
        if (rayMinT > mint) mint = rayMinT;
        if (ray.maxt < maxt) maxt = ray.maxt;
        if (!(maxt > mint)) return;
        RayStatistics st = rayIntersectHavranCollectStatistics(ray, mint, maxt, t, temp);
        c[0] += st.numTraversals;
        c[1] += st.numIntersections;
        c[2] += st.foundIntersection ? 1 : 0;
    }
};
int ref_kd_count(void *s, const float *rays, size_t n, uint64_t *counters) {
    REF_TRY
    RefScene *rs = (RefScene *)s;
    ensureBuilt(rs);
    const KDProbe *probe = static_cast<const KDProbe *>(rs->scene->getKDTree());
    counters[0] = counters[1] = counters[2] = 0;
    for (size_t i = 0; i < n; ++i) {
        const float *r = rays + 8 * i;
        probe->count(Ray(Point(r[0], r[1], r[2]), Vector(r[4], r[5], r[6]), r[3], r[7], 0.0f), counters);
    }
    return 0;
    REF_CATCH(-1)
}

// Depth-first dump of the reference's kd-tree: per node 3 floats {axis (or -1 for a leaf), split (or primitive count), depth}
int ref_kd_dump(void *s, float *out, int max_nodes) {
    REF_TRY
    RefScene *rs = (RefScene *)s;
    ensureBuilt(rs);
    const ShapeKDTree *tree = rs->scene->getKDTree();
    typedef ShapeKDTree::KDNode Node;
    std::vector<std::pair<const Node *, int>> stack;
    stack.push_back(std::make_pair(tree->getRoot(), 0));
    int n = 0;
    while (!stack.empty() && n < max_nodes) {
        const Node *node = stack.back().first;
        int depth = stack.back().second;
        stack.pop_back();
        if (node->isLeaf()) {
            out[3 * n] = -1; out[3 * n + 1] = (float)(node->getPrimEnd() - node->getPrimStart()); out[3 * n + 2] = (float)depth;
        } else {
            out[3 * n] = (float)node->getAxis(); out[3 * n + 1] = node->getSplit(); out[3 * n + 2] = (float)depth;
            stack.push_back(std::make_pair(node->getRight(), depth + 1));
            stack.push_back(std::make_pair(node->getLeft(), depth + 1));
        }
        ++n;
    }
    return n;
    REF_CATCH(-1)
}

// A `heterogeneous` medium the way the scene loader would assemble it from XML: plugin + named properties for the medium, its
// `density` gridvolume (reads the .vol FILE named in its properties -- the reference's own reader, gridvolume.cpp:218-290), its
// `albedo` constvolume and its phase function. Returns a scene handle whose medium 0 is that medium (for ref_grid_lookup,
// ref_medium_sample, ref_phase); seed for the replay stream.
void *ref_medium_from_props(const char *medium_props, const char *density_props, const char *albedo_props, const char *phase_plugin,
                            const char *phase_props, uint64_t seed) {
    REF_TRY
    ensureInit();
    std::unique_ptr<RefScene> rs(new RefScene());
    rs->seed = seed;
    rs->width = 1 << 20;
    ref<Medium> med = create<Medium>(parseProps("heterogeneous", medium_props));
    ref<VolumeDataSource> dens = create<VolumeDataSource>(parseProps("gridvolume", density_props));
    dens->configure();
    ref<VolumeDataSource> alb = create<VolumeDataSource>(parseProps("constvolume", albedo_props));
    alb->configure();
    med->addChild("density", dens);
    med->addChild("albedo", alb);
    if (phase_plugin && *phase_plugin) {
        ref<PhaseFunction> phase = create<PhaseFunction>(parseProps(phase_plugin, phase_props));
        phase->configure();
        med->addChild(phase);
    }
    med->configure();
    rs->media.push_back(med);
    rs->densities.push_back(dens);
    return rs.release();
    REF_CATCH(nullptr)
}

int ref_num_threads() { return omp_get_max_threads(); }

}  // extern "C"
