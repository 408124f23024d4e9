// TEST INFRASTRUCTURE ONLY -- see oracle_math.h header note.
//
// oracle_medium.h: heterogeneous medium over a dense grid volume, restated from
//   src/medium/heterogeneous.cpp:546-663 (Woodcock tracking), src/volume/gridvolume.cpp:188-215, 337-388,
//   src/phase/hg.cpp:74-110, src/phase/isotropic.cpp:62-78.
#pragma once
#include <vector>

#include "../include/b200pg.h"
#include "oracle_math.h"

namespace orc {

struct Medium {
    B200pgMedium d;
    std::vector<Float> density;
    static Medium fromDesc(const B200pgMedium &m) {
        Medium r;
        r.d = m;
        size_t n = (size_t)m.res[0] * m.res[1] * m.res[2];
        if (m.density) r.density.assign(m.density, m.density + n);
        r.d.density = nullptr;
        return r;
    }
};

}  // namespace orc
