"""Writes the benchmark scenes (BASELINE.json configs C1-C4) as ordinary Mitsuba 0.6 XML (+ .serialized / .vol
files) so that they can be rendered with `b200pg-render` -- or with the reference itself elsewhere.
usage: python tools/export_scenes.py out_dir [--small]      (--small: reduced resolutions / mesh size for a quick look)"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge  # noqa: E402


def main():
    if len(sys.argv) < 2:
        raise SystemExit(__doc__)
    out, small = sys.argv[1], "--small" in sys.argv[2:]
    S = ge.load_package().scenes
    guided = dict(type="guidedpath", maxDepth=8, trainingProgressions=16, samplesPerProgression=4, maxComponents=16,
                  maxSamplesPerCell=32768)
    cases = {
        "c1_cornell": S.cornell_box(128 if small else 512, 128 if small else 512, spp=64),
        "c2_cornell_caustic_guided": S.cornell_caustic(256 if small else 1024, 256 if small else 1024, spp=64),
        "c3_medium_guided": S.cornell_medium(256 if small else 1024, 256 if small else 1024, spp=64, res=32 if small else 256),
        "c4_mesh_guided": S.mesh_scene(256 if small else 2048, 256 if small else 2048, spp=16, **(dict(n=129) if small else {})),
    }
    cases["c2_cornell_caustic_guided"].integrator = dict(guided)
    cases["c3_medium_guided"].integrator = dict(type="guidedvolpath", maxDepth=8, trainingProgressions=16, samplesPerProgression=4,
                                                guidedDistanceSampling=True)
    cases["c4_mesh_guided"].integrator = dict(guided)
    for name, sb in cases.items():
        d = os.path.join(out, name)
        os.makedirs(d, exist_ok=True)
        print(S.save_scene(sb, d))


if __name__ == "__main__":
    main()
