// b200guidedpath.cpp -- the REFERENCE-SIDE binding of libb200pg.so: a Mitsuba integrator plugin, written against the
// reference's own plugin interface, that a maintainer drops into src/integrators/path/ (add_integrator(b200guidedpath ...)
// in src/integrators/CMakeLists.txt:14-52, link -lb200pg). It forwards the reference's virtual calls to the C-ABI of
// include/b200pg.h and hands the result back through the reference's own Film:
//
//   MTS_EXPORT_PLUGIN / CreateInstance(const Properties &)      parameters -> B200pgIntegratorParams       (cobject.h:99-107)
//   Integrator::preprocess                                       nothing to do (no per-pixel samplers)       (integrator.h:61)
//   Integrator::render                                           b200pg_scene_load_xml + b200pg_integrator_create +
//                                                                b200pg_render, then Film::put of the GPU film (integrator.h:74)
//   Integrator::cancel                                           b200pg_cancel (async-safe)                  (integrator.h:84)
//   Integrator::postprocess                                      the statistics the reference logs           (progressive_path.cpp:124-130)
//
// Parameter names are those of `progressivepath` / `progressivevolpath` (integrator.cpp:195-230,
// progressiveintegrator.cpp:296-300, progressive_path.cpp:117); the guiding parameters are this library's.
// Here it is compiled by integration/Makefile against the reference build of oracle/Makefile.ref (oracle/_ref) and loaded by
// the reference's PluginManager in tests/test_zz_integration_plugin.py.
#include <mitsuba/render/scene.h>
#include <mitsuba/render/film.h>
#include <mitsuba/render/imageblock.h>
#include <mitsuba/render/progressiveintegrator.h>
#include <mitsuba/core/statistics.h>
#include <mitsuba/core/fresolver.h>  // boost::filesystem::path (Scene::getSourceFile)
#include "b200pg.h"

MTS_NAMESPACE_BEGIN

/// The GPU film (R, G, B, alpha, weight per pixel, row-major, crop-sized: ImageBlock's ESpectrumAlphaWeight layout,
/// imageblock.h:131-138) goes into the reference's film the way a worker's image block does (hdrfilm.cpp:391-393).
/// extern "C" so that the test harness can drive this one step without a device.
extern "C" void b200guidedpath_put_film(Film *film, const float *rgbaw) {
    const Vector2i size = film->getCropSize();
    ref<ImageBlock> block = new ImageBlock(Bitmap::ESpectrumAlphaWeight, size, film->getReconstructionFilter());
    block->setOffset(film->getCropOffset());
    block->clear();
    const int border = block->getBorderSize(), stride = size.x + 2 * border;
    Float *dst = block->getBitmap()->getFloatData();
    for (int y = 0; y < size.y; ++y)
        memcpy(dst + ((size_t) (y + border) * stride + border) * 5, rgbaw + (size_t) y * size.x * 5, sizeof(float) * 5 * size.x);
    film->put(block);
}

class B200GuidedPathTracer : public ProgressiveMonteCarloIntegrator {
public:
    B200GuidedPathTracer(const Properties &props) : ProgressiveMonteCarloIntegrator(props), m_handle(NULL) {
        b200pg_integrator_params_default(&m_p);
        m_p.max_depth = m_maxDepth;  // integrator.cpp:195-230
        m_p.rr_depth = m_rrDepth;
        m_p.strict_normals = m_strictNormals;
        m_p.hide_emitters = m_hideEmitters;
        m_p.samples_per_progression = m_samplesPerProgression;  // progressiveintegrator.cpp:296-300
        m_p.max_render_time = (int) m_maxRenderTime;
        m_p.max_component_value = m_maxComponentValue;
        m_p.use_nee = props.getBoolean("useNee", true);  // progressive_path.cpp:117
        m_p.volumetric = props.getBoolean("volumetric", props.getPluginName() == "b200guidedvolpath");
        m_p.guiding = props.getBoolean("guiding", true);
        m_p.training_progressions = props.getInteger("trainingProgressions", m_p.training_progressions);
        m_p.guiding_probability = props.getFloat("guidingProbability", m_p.guiding_probability);
        m_p.guide_max_components = props.getInteger("maxComponents", m_p.guide_max_components);
        m_p.guide_max_cell_samples = props.getInteger("maxSamplesPerCell", m_p.guide_max_cell_samples);
        m_p.guide_train_discard_film = props.getBoolean("discardTrainingSamples", m_p.guide_train_discard_film != 0);
        m_p.guided_distance = props.getBoolean("guidedDistance", false);
        m_device = props.getInteger("device", 0);
        m_deviceCount = props.getInteger("deviceCount", 1);  // GPUs of this box used inside one render() call
    }

    B200GuidedPathTracer(Stream *stream, InstanceManager *manager)
        : ProgressiveMonteCarloIntegrator(stream, manager), m_handle(NULL), m_device(0), m_deviceCount(1) {
        b200pg_integrator_params_default(&m_p);
    }

    /// The per-pixel sampler table of ProgressiveMonteCarloIntegrator::preprocess is not needed: the device derives
    /// every random number from (seed, pixel, sample index)
    bool preprocess(const Scene *, RenderQueue *, const RenderJob *, int, int, int) { return true; }

    bool render(Scene *scene, RenderQueue *queue, const RenderJob *job, int, int, int) {
        const std::string source = scene->getSourceFile().string();
        if (source.empty())
            Log(EError, "b200guidedpath: the scene has no source file (Scene::setSourceFile); the GPU library reads the same "
                "scene XML the reference was given");
        char err[1024] = "";
        void *gpuScene = b200pg_scene_load_xml(source.c_str(), NULL, err, sizeof(err));
        if (!gpuScene)
            Log(EError, "b200guidedpath: %s", err);  // EError throws (renderjob.cpp:111-115)
        const B200pgSceneDesc *desc = b200pg_scene_desc(gpuScene);
        Film *film = scene->getFilm();
        if (desc->film.width != film->getCropSize().x || desc->film.height != film->getCropSize().y) {
            b200pg_scene_destroy(gpuScene);
            Log(EError, "b200guidedpath: film size of the XML (%ix%i) and of the live scene (%ix%i) differ",
                desc->film.width, desc->film.height, film->getCropSize().x, film->getCropSize().y);
        }
        void *h = b200pg_integrator_create(gpuScene, &m_p, m_device);
        if (!h) {
            std::string why = b200pg_last_error();
            b200pg_scene_destroy(gpuScene);
            Log(EError, "b200guidedpath: %s", why.c_str());
        }
        m_handle = h;
        std::vector<int> devices(std::max(m_deviceCount, 1));
        devices[0] = m_device;  // devices[0] is the integrator's own device (b200pg.h)
        for (size_t i = 1, d = 0; i < devices.size(); ++d)
            if ((int) d != m_device) devices[i++] = (int) d;
        Log(EInfo, "Starting render job on %i B200 device(s) (%ix%i, " SIZE_T_FMT " samples)", (int) devices.size(),
            film->getCropSize().x, film->getCropSize().y, scene->getSampler()->getSampleCount());
        const int rc = b200pg_render(h, (int) devices.size(), devices.data());  // blocking, like Integrator::render
        std::string why = rc ? b200pg_last_error() : "";
        if (rc == 0) {
            const Vector2i size = film->getCropSize();
            std::vector<float> rgbaw((size_t) size.x * size.y * 5);
            b200pg_film_read(h, rgbaw.data());
            b200guidedpath_put_film(film, rgbaw.data());
            b200pg_stats(h, &m_stats);
            m_spp = (int) (m_stats.paths / std::max((uint64_t) 1, (uint64_t) size.x * size.y));
        }
        m_handle = NULL;
        b200pg_destroy(h);
        b200pg_scene_destroy(gpuScene);
        if (rc != 0)
            Log(EError, "b200guidedpath: %s", why.c_str());
        return true;
    }

    void cancel() {
        void *h = m_handle;
        if (h) b200pg_cancel(h);
    }

    void postprocess(const Scene *, RenderQueue *, const RenderJob *, int, int, int) {
        Log(EInfo, "Rendered samples: %d; normal rays traced: " SIZE_T_FMT ", shadow rays traced: " SIZE_T_FMT
            ", avg. path length: %f, guiding cells: %u", m_spp, (size_t) m_stats.normal_rays, (size_t) m_stats.shadow_rays,
            m_stats.paths ? (double) m_stats.path_length_sum / (double) m_stats.paths : 0.0, m_stats.guide_cells);
    }

    /// Not used: radiance is estimated on the device
    Spectrum Li(const RayDifferential &, RadianceQueryRecord &) const { return Spectrum(0.0f); }

    std::string toString() const {
        std::ostringstream oss;
        oss << "B200GuidedPathTracer[" << endl << "  maxDepth = " << m_p.max_depth << "," << endl << "  rrDepth = " << m_p.rr_depth << ","
            << endl << "  guiding = " << m_p.guiding << "," << endl << "  trainingProgressions = " << m_p.training_progressions << ","
            << endl << "  maxComponents = " << m_p.guide_max_components << "," << endl << "  volumetric = " << m_p.volumetric << ","
            << endl << "  deviceCount = " << m_deviceCount << endl << "]";
        return oss.str();
    }

    MTS_DECLARE_CLASS()
private:
    B200pgIntegratorParams m_p;
    B200pgStats m_stats = {};
    void *volatile m_handle;
    int m_device, m_deviceCount;
};

MTS_IMPLEMENT_CLASS_S(B200GuidedPathTracer, false, ProgressiveMonteCarloIntegrator)
MTS_EXPORT_PLUGIN(B200GuidedPathTracer, "B200 guided path tracer");
MTS_NAMESPACE_END
