// b200pg-render -- command-line host mirroring the subset of the `mitsuba` CLI that matters on this path
// (src/mitsuba/mitsuba.cpp:52-91): -o <file>, -D key=value, -p <count>, -r <sec>, -q, -v. Uses only the C-ABI.
//
// -p n (mitsuba: number of local worker threads) = number of GPUs. n > 1 runs ONE PROCESS PER GPU: the launcher forks
// n workers before anything touches CUDA and only carries 128-byte IPC handles and go / stop bytes over pipes. Every
// worker loads the scene, renders its own sample batches (global pass g: worker r renders sample block g*n + r), sums the
// guiding field's EM statistics with its peers inside the M-step kernel over NVLink peer memory (b200pg_comm_connect), and
// worker 0 finally adds the peers' films into its own (b200pg_film_add_peers) and writes the image -- the job
// ProgressiveMonteCarloIntegrator::renderSamples / renderTime does with its BlockedRenderProcess workers
// (progressiveintegrator.cpp:65-168), minus the network layer.
#include <signal.h>
#include <sys/types.h>
#include <sys/wait.h>
#include <unistd.h>

#include <chrono>
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
                "   -p count     Number of GPUs to render on, one worker process each (default 1)\n"
                "   -d index     First CUDA device (default 0; worker r uses device index + r)\n"
                "   -r sec       Render for a time budget instead of the sampler's sample count (maxRenderTime)\n"
                "   -q           Quiet mode\n"
                "   -v           Verbose: one line per progression\n");
}

struct Options {
    std::string out, scenePath;
    std::vector<std::string> defs;
    int gpus = 1, device = 0, seconds = -1;
    bool quiet = false, verbose = false;
};

static bool writeAll(int fd, const void *p, size_t n) {
    const char *c = (const char *)p;
    while (n) {
        ssize_t k = write(fd, c, n);
        if (k <= 0) return false;
        c += k;
        n -= (size_t)k;
    }
    return true;
}
static bool readAll(int fd, void *p, size_t n) {
    char *c = (char *)p;
    while (n) {
        ssize_t k = read(fd, c, n);
        if (k <= 0) return false;
        c += k;
        n -= (size_t)k;
    }
    return true;
}

static std::string outputName(const Options &o, void *scene) {
    if (!o.out.empty()) return o.out;
    std::string out = o.scenePath;
    size_t dot = out.find_last_of('.');
    if (dot != std::string::npos) out = out.substr(0, dot);
    const B200pgSceneDesc *desc = b200pg_scene_desc(scene);
    const int ff = desc ? desc->film.file_format : 0;
    return out + (ff == 1 ? ".pfm" : (ff == 2 ? ".rgbe" : ".exr"));
}

static void report(const B200pgStats &st, const std::string &out, int gpus) {
    // the reference logs "Render time" (renderjob.cpp:108), ray counters (skdtree.cpp:46-47) and the average path length
    std::printf("Render time: %.4fs (device%s)\n", st.seconds_total, gpus > 1 ? ", slowest GPU" : "");
    std::printf("Normal rays traced: %llu\nShadow rays traced: %llu\n", (unsigned long long)st.normal_rays, (unsigned long long)st.shadow_rays);
    std::printf("Avg. path length: %f (%llu/%llu)\n", st.paths ? (double)st.path_length_sum / st.paths : 0.0,
                (unsigned long long)st.path_length_sum, (unsigned long long)st.paths);
    std::printf("Guiding cells: %u, training samples: %llu\nWrote %s\n", st.guide_cells, (unsigned long long)st.train_samples, out.c_str());
}

static void *loadScene(const Options &o) {
    std::vector<const char *> dptr;
    for (auto &d : o.defs) dptr.push_back(d.c_str());
    dptr.push_back(nullptr);
    char err[1024] = {0};
    void *scene = b200pg_scene_load_xml(o.scenePath.c_str(), dptr.data(), err, sizeof(err));
    if (!scene) std::fprintf(stderr, "Error: %s\n", err);
    return scene;
}

// ---- one GPU: the library's own progression loop -------------------------------------------------------------------
static int renderSingle(const Options &o) {
    void *scene = loadScene(o);
    if (!scene) return 2;
    B200pgIntegratorParams p;
    b200pg_scene_integrator_params(scene, &p);
    if (o.seconds >= 0) p.max_render_time = o.seconds;
    void *integ = b200pg_integrator_create(scene, &p, o.device);
    if (!integ) { std::fprintf(stderr, "Error: %s\n", b200pg_last_error()); return 2; }
    if (b200pg_render(integ, 1, &o.device) != 0) { std::fprintf(stderr, "Error: %s\n", b200pg_last_error()); return 3; }
    const std::string out = outputName(o, scene);
    if (b200pg_film_write(integ, out.c_str()) != 0) { std::fprintf(stderr, "Error: %s\n", b200pg_last_error()); return 4; }
    B200pgStats st;
    b200pg_stats(integ, &st);
    if (!o.quiet) report(st, out, 1);
    b200pg_destroy(integ);
    b200pg_scene_destroy(scene);
    return 0;
}

// ---- n GPUs: worker process of rank `rank`; `up` = pipe to the launcher, `down` = pipe from it -----------------------
struct Handles {
    unsigned char comm[64], film[64];
};
enum : char { kGo = 'g', kStop = 's' };

static int worker(const Options &o, int rank, int world, int up, int down) {
#define PG_CHECK(call, code) \
    if ((call) != 0) { std::fprintf(stderr, "Error (GPU %d): %s\n", o.device + rank, b200pg_last_error()); return code; }
    void *scene = loadScene(o);
    if (!scene) return 2;
    B200pgIntegratorParams p;
    b200pg_scene_integrator_params(scene, &p);
    if (o.seconds >= 0) p.max_render_time = o.seconds;
    void *integ = b200pg_integrator_create(scene, &p, o.device + rank);
    if (!integ) { std::fprintf(stderr, "Error (GPU %d): %s\n", o.device + rank, b200pg_last_error()); return 2; }
    // exchange the IPC handles through the launcher
    Handles mine;
    std::memset(&mine, 0, sizeof(mine));
    if (p.guiding) PG_CHECK(b200pg_comm_local_handle(integ, mine.comm), 2);
    PG_CHECK(b200pg_film_ipc_handle(integ, mine.film), 2);
    std::vector<Handles> all((size_t)world);
    if (!writeAll(up, &mine, sizeof(mine)) || !readAll(down, all.data(), sizeof(Handles) * (size_t)world)) return 5;
    std::vector<unsigned char> comm((size_t)world * 64), film((size_t)world * 64);
    for (int r = 0; r < world; ++r) {
        std::memcpy(&comm[64 * (size_t)r], all[(size_t)r].comm, 64);
        std::memcpy(&film[64 * (size_t)r], all[(size_t)r].film, 64);
    }
    if (p.guiding) PG_CHECK(b200pg_comm_connect(integ, rank, world, comm.data()), 2);

    // progression loop (progressiveintegrator.cpp:65-168): global pass g = `world` sample blocks, one per worker
    const B200pgSceneDesc *desc = b200pg_scene_desc(scene);
    const int perPass = p.samples_per_progression > 0 ? p.samples_per_progression : 1;
    const int numPasses = std::max(1, (desc ? desc->sample_count : 4) / perPass);
    const int globalPasses = (numPasses + world - 1) / world;  // the sample count is rounded up to a multiple of world * perPass
    const bool timed = p.max_render_time > 0;
    // global passes that train the field: as b200pg_render counts them (integrator.cu: makePlan) -- the single-device number of
    // training SAMPLES under a sample budget, training_progressions UPDATES under a time budget
    const int T = p.guiding ? std::max(0, p.training_progressions) : 0;
    const int trainPasses = timed ? T : std::min((T + world - 1) / world, globalPasses);
    const auto t0 = std::chrono::steady_clock::now();
    for (int g = 0; timed || g < globalPasses; ++g) {
        const bool record = g < trainPasses;
        if (p.guiding) PG_CHECK(b200pg_guiding_mode(integ, record ? 1 : 0, 1), 3);
        PG_CHECK(b200pg_progression_render(integ, (g * world + rank) * perPass, perPass, 0, 0), 3);
        if (record) {
            uint32_t ns = 0, nc = 0;
            PG_CHECK(b200pg_train(integ, 0, &ns, &nc), 3);  // statistics summed over all workers inside the M-step kernel
            if (o.verbose && rank == 0) std::printf("Progression[%d]: %u local training samples, %u cells\n", g, ns, nc);
            if (p.guide_train_discard_film && g + 1 >= trainPasses) PG_CHECK(b200pg_film_clear(integ), 3);
        } else if (o.verbose && rank == 0) {
            std::printf("Progression[%d] took %.3f s so far\n", g, std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count());
        }
        if (timed) {  // the launcher decides for everybody (worker 0's clock), so that all workers leave the loop together
            const double el = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
            char verdict = kStop;
            if (!writeAll(up, &el, sizeof(el)) || !readAll(down, &verdict, 1)) return 5;
            if (verdict != kGo) break;
        }
    }
    // done: statistics to the launcher, then wait until worker 0 has merged the films
    B200pgStats st;
    b200pg_stats(integ, &st);
    if (!writeAll(up, &st, sizeof(st))) return 5;
    char c = 0;
    if (!readAll(down, &c, 1)) return 5;  // rank 0: "all workers are done"; others: "worker 0 has read your film"
    int rc = 0;
    if (rank == 0) {
        const std::string out = outputName(o, scene);
        if (b200pg_film_add_peers(integ, rank, world, film.data()) != 0 || b200pg_film_write(integ, out.c_str()) != 0) {
            std::fprintf(stderr, "Error: %s\n", b200pg_last_error());
            rc = 4;
        }
        char ok = rc == 0 ? kGo : kStop;
        if (!writeAll(up, &ok, 1)) rc = 5;
    }
    b200pg_destroy(integ);
    b200pg_scene_destroy(scene);
    return rc;
#undef PG_CHECK
}

static int renderMulti(const Options &o) {
    const int world = o.gpus;
    std::vector<pid_t> pids((size_t)world, -1);
    std::vector<int> fromChild((size_t)world, -1), toChild((size_t)world, -1);
    signal(SIGPIPE, SIG_IGN);
    std::fflush(stdout);
    std::fflush(stderr);
    for (int r = 0; r < world; ++r) {
        int up[2], down[2];
        if (pipe(up) != 0 || pipe(down) != 0) { std::perror("pipe"); return 5; }
        const pid_t pid = fork();  // before anything in this process has touched CUDA
        if (pid < 0) { std::perror("fork"); return 5; }
        if (pid == 0) {
            close(up[0]);
            close(down[1]);
            for (int q = 0; q < r; ++q) { close(fromChild[(size_t)q]); close(toChild[(size_t)q]); }
            const int rc = worker(o, r, world, up[1], down[0]);
            std::fflush(stdout);
            std::fflush(stderr);
            _exit(rc);
        }
        close(up[1]);
        close(down[0]);
        pids[(size_t)r] = pid;
        fromChild[(size_t)r] = up[0];
        toChild[(size_t)r] = down[1];
    }
    auto abortAll = [&](const char *why) {
        std::fprintf(stderr, "Error: %s\n", why);
        for (int r = 0; r < world; ++r) { close(toChild[(size_t)r]); close(fromChild[(size_t)r]); }
        int worst = 5;
        for (int r = 0; r < world; ++r) {  // closed pipes make every worker's next read / write fail; reap them
            int status = 0;
            if (waitpid(pids[(size_t)r], &status, 0) > 0 && WIFEXITED(status) && WEXITSTATUS(status) != 0 && WEXITSTATUS(status) != 5)
                worst = WEXITSTATUS(status);
        }
        return worst;
    };
    // 1. gather and broadcast the IPC handles
    std::vector<Handles> all((size_t)world);
    for (int r = 0; r < world; ++r)
        if (!readAll(fromChild[(size_t)r], &all[(size_t)r], sizeof(Handles))) return abortAll("a worker failed during start-up");
    for (int r = 0; r < world; ++r)
        if (!writeAll(toChild[(size_t)r], all.data(), sizeof(Handles) * (size_t)world)) return abortAll("a worker failed during start-up");
    // 2. time-budget mode: one verdict per global pass for everybody. The launcher parses the scene as well (host only, no
    // CUDA) to learn the integrator's maxRenderTime and the default output name.
    bool timed = o.seconds > 0;
    double limit = o.seconds > 0 ? (double)o.seconds : 0.0;
    std::string outName = o.out;
    {
        void *scene = loadScene(o);
        if (!scene) return abortAll("cannot load the scene");
        B200pgIntegratorParams p;
        b200pg_scene_integrator_params(scene, &p);
        if (o.seconds < 0 && p.max_render_time > 0) {
            timed = true;
            limit = p.max_render_time;
        }
        outName = outputName(o, scene);
        b200pg_scene_destroy(scene);
    }
    while (timed) {
        double el0 = 0;
        for (int r = 0; r < world; ++r) {
            double el = 0;
            if (!readAll(fromChild[(size_t)r], &el, sizeof(el))) return abortAll("a worker failed while rendering");
            if (r == 0) el0 = el;
        }
        const char verdict = el0 >= limit ? kStop : kGo;
        for (int r = 0; r < world; ++r)
            if (!writeAll(toChild[(size_t)r], &verdict, 1)) return abortAll("a worker failed while rendering");
        if (verdict == kStop) break;
    }
    // 3. statistics of every worker = "this worker has finished rendering"
    B200pgStats sum;
    std::memset(&sum, 0, sizeof(sum));
    for (int r = 0; r < world; ++r) {
        B200pgStats st;
        if (!readAll(fromChild[(size_t)r], &st, sizeof(st))) return abortAll("a worker failed while rendering");
        sum.paths += st.paths;
        sum.normal_rays += st.normal_rays;
        sum.shadow_rays += st.shadow_rays;
        sum.path_length_sum += st.path_length_sum;
        sum.kernel_launches += st.kernel_launches;
        sum.train_samples += st.train_samples;
        sum.seconds_total = std::max(sum.seconds_total, st.seconds_total);
        sum.guide_cells = st.guide_cells;
        sum.progressions_done = std::max(sum.progressions_done, st.progressions_done);
    }
    // 4. worker 0 merges the films while the others keep theirs alive, then everybody may exit
    const char go = kGo;
    char merged = kStop;
    if (!writeAll(toChild[0], &go, 1) || !readAll(fromChild[0], &merged, 1)) return abortAll("worker 0 failed while merging the films");
    for (int r = 1; r < world; ++r) writeAll(toChild[(size_t)r], &go, 1);
    int rc = merged == kGo ? 0 : 4;
    for (int r = 0; r < world; ++r) {
        int status = 0;
        waitpid(pids[(size_t)r], &status, 0);
        if (rc == 0 && !(WIFEXITED(status) && WEXITSTATUS(status) == 0)) rc = WIFEXITED(status) ? WEXITSTATUS(status) : 5;
    }
    if (rc == 0 && !o.quiet) {
        std::printf("GPUs: %d (one worker process each)\n", world);
        report(sum, outName, world);
    }
    return rc;
}

int main(int argc, char **argv) {
    Options o;
    for (int i = 1; i < argc; ++i) {
        std::string a = argv[i];
        if (a == "-o" && i + 1 < argc) o.out = argv[++i];
        else if (a == "-D" && i + 1 < argc) o.defs.push_back(argv[++i]);
        else if (a.rfind("-D", 0) == 0 && a.size() > 2) o.defs.push_back(a.substr(2));
        else if (a == "-p" && i + 1 < argc) o.gpus = std::atoi(argv[++i]);
        else if (a == "-d" && i + 1 < argc) o.device = std::atoi(argv[++i]);
        else if (a == "-r" && i + 1 < argc) o.seconds = std::atoi(argv[++i]);
        else if (a == "-q") o.quiet = true;
        else if (a == "-v") o.verbose = true;
        else if (a == "-h") { usage(); return 0; }
        else if (a[0] == '-') { std::fprintf(stderr, "unknown option %s\n", a.c_str()); usage(); return 1; }
        else o.scenePath = a;
    }
    if (o.scenePath.empty()) { usage(); return 1; }
    if (o.gpus < 1 || o.gpus > 16) { std::fprintf(stderr, "Error: -p expects a GPU count between 1 and 16\n"); return 1; }
    if (o.device < 0) { std::fprintf(stderr, "Error: -d expects a CUDA device index\n"); return 1; }
    return o.gpus == 1 ? renderSingle(o) : renderMulti(o);
}
