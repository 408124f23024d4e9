// TEST INFRASTRUCTURE ONLY -- see oracle_math.h header note.
//
// oracle_volpath.h: ProgressiveVolumetricPathTracer::Li (progressive_volpath.cpp:98-460).
#pragma once

namespace orc {

static Vec3 Li_volpath(const Scene &scene, const B200pgIntegratorParams &P, const Ray &r, Rng &rng, Stats &st) {
    return Li_path(scene, P, r, rng, st);  // filled in with the medium row
}

}  // namespace orc
