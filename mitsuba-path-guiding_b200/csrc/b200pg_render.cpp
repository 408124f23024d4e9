// b200pg-render -- command-line host mirroring the subset of the `mitsuba` CLI that matters on this path
// (src/mitsuba/mitsuba.cpp:52-91): -o <file>, -D key=value, -p <gpu index>, -q. Uses only the C-ABI.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/b200pg.h"

static void usage() {
    std::printf("Usage: b200pg-render [options] <scene.xml>\n"
                "   -o fname     Write the developed image to fname (.exr, .pfm or .rgbe). Default: <scene> + the extension of\n"
                "                the film's fileFormat (openexr unless the scene says otherwise, hdrfilm.cpp:212-226)\n"
                "   -D key=val   Define a constant, which can be referenced as \"$key\" in the scene\n"
                "   -p index     CUDA device to render on (default 0)\n"
                "   -q           Quiet mode\n");
}

int main(int argc, char **argv) {
    std::string out, scenePath;
    std::vector<std::string> defs;
    int device = 0;
    bool quiet = false;
    for (int i = 1; i < argc; ++i) {
        std::string a = argv[i];
        if (a == "-o" && i + 1 < argc) out = argv[++i];
        else if (a == "-D" && i + 1 < argc) defs.push_back(argv[++i]);
        else if (a.rfind("-D", 0) == 0 && a.size() > 2) defs.push_back(a.substr(2));
        else if (a == "-p" && i + 1 < argc) device = std::atoi(argv[++i]);
        else if (a == "-q") quiet = true;
        else if (a == "-h") { usage(); return 0; }
        else if (a[0] == '-') { std::fprintf(stderr, "unknown option %s\n", a.c_str()); usage(); return 1; }
        else scenePath = a;
    }
    if (scenePath.empty()) { usage(); return 1; }
    std::vector<const char *> dptr;
    for (auto &d : defs) dptr.push_back(d.c_str());
    dptr.push_back(nullptr);
    char err[1024] = {0};
    void *scene = b200pg_scene_load_xml(scenePath.c_str(), dptr.data(), err, sizeof(err));
    if (!scene) { std::fprintf(stderr, "Error: %s\n", err); return 2; }
    B200pgIntegratorParams p;
    b200pg_scene_integrator_params(scene, &p);
    void *integ = b200pg_integrator_create(scene, &p, device);
    if (!integ) { std::fprintf(stderr, "Error: %s\n", b200pg_last_error()); return 2; }
    if (b200pg_render(integ) != 0) { std::fprintf(stderr, "Error: %s\n", b200pg_last_error()); return 3; }
    if (out.empty()) {
        out = scenePath;
        size_t dot = out.find_last_of('.');
        if (dot != std::string::npos) out = out.substr(0, dot);
        const B200pgSceneDesc *desc = b200pg_scene_desc(scene);
        const int ff = desc ? desc->film.file_format : 0;
        out += ff == 1 ? ".pfm" : (ff == 2 ? ".rgbe" : ".exr");
    }
    if (b200pg_film_write(integ, out.c_str()) != 0) { std::fprintf(stderr, "Error: %s\n", b200pg_last_error()); return 4; }
    B200pgStats st;
    b200pg_stats(integ, &st);
    if (!quiet) {
        // the reference logs "Render time" (renderjob.cpp:108), ray counters (skdtree.cpp:46-47) and the average path length
        std::printf("Render time: %.4fs (device)\n", st.seconds_total);
        std::printf("Normal rays traced: %llu\nShadow rays traced: %llu\n", (unsigned long long)st.normal_rays, (unsigned long long)st.shadow_rays);
        std::printf("Avg. path length: %f (%llu/%llu)\n", st.paths ? (double)st.path_length_sum / st.paths : 0.0,
                    (unsigned long long)st.path_length_sum, (unsigned long long)st.paths);
        std::printf("Guiding cells: %u, training samples: %llu\nWrote %s\n", st.guide_cells, (unsigned long long)st.train_samples, out.c_str());
    }
    b200pg_destroy(integ);
    b200pg_scene_destroy(scene);
    return 0;
}
