// host_scene.h -- host-side scene compiler: owns a deep copy of the flat description
// (include/b200pg.h), assigns default BSDFs like Shape::configure (src/librender/shape.cpp:48-70),
// numbers primitives like ShapeKDTree (src/librender/skdtree.cpp:53-104), builds the BVH and
// emits the GPU records of pg_types.h.
#pragma once
#include <string>
#include <vector>

#include "../../include/b200pg.h"
#include "pg_types.h"

namespace pg {

struct HostScene {
    // ---- deep copy of the description
    std::vector<B200pgShape> shapes;
    std::vector<B200pgBsdf> bsdfs;
    std::vector<B200pgEmitter> emitters;
    std::vector<B200pgMedium> media;
    std::vector<std::vector<float>> ownedF;      // positions / normals / texcoords / densities
    std::vector<std::vector<uint32_t>> ownedU;   // indices
    B200pgSensor sensor;
    B200pgFilm film;
    int sampleCount = 4;
    uint64_t seed = 1337;
    B200pgIntegratorParams xmlParams;  // integrator found in the XML (defaults otherwise)
    B200pgSceneDesc view;              // borrowed view handed out by b200pg_scene_desc

    // ---- compiled records
    std::vector<BvhNode> nodes;
    std::vector<WideNode> wideNodes;     // 8-wide quantised tree over the same leaves (large meshes only; empty otherwise)
    int wideDepth = 0;
    // primitive records in BVH leaf order, one per SLOT (leaves start at even slots; holes are never referenced):
    std::vector<float> primPlanes;       // 4 floats: row 2 of the affine map (pg_types.h: PrimRecord), the plane distance
    std::vector<float> primRows;         // 8 floats: rows 0 / 1, the (u, v) coordinates
    size_t nPrimitives = 0;              // primitives (slots minus holes)
    std::vector<uint32_t> primGlobalId;  // BVH order -> global primitive id
    std::vector<PrimInfo> primInfo;      // BVH order -> (shape, primitive index)
    std::vector<float> shadeTris;        // BVH order -> ShadeTri (24 floats), what the shade stage reads per hit
    std::vector<RectRecord> rects;
    std::vector<ShapeRecord> shapeRecs;
    std::vector<MeshRecord> meshes;
    std::vector<float> positions, normals, texcoords;  // concatenated pools (3,3,2 floats / vertex)
    std::vector<uint32_t> indices;                     // 3 / triangle
    std::vector<float> areaCdf;
    std::vector<BsdfRecord> bsdfRecs;
    std::vector<EmitterRecord> emitterRecs;
    std::vector<float> emitterCdf;
    std::vector<MediumRecord> mediumRecs;
    std::vector<float> densityPool;
    CameraRecord camera;
    FilmRecord filmRec;
    float sceneMin[3], sceneMax[3];
    uint32_t primCount = 0;
    int rootIsLeaf = 0;

    bool copyFrom(const B200pgSceneDesc *desc, std::string &err);
    bool compile(std::string &err);
    void refreshView();
};

// Rough-transmittance tables (packed layout of tools/pack_rtrans.py); reduction as rtrans.h:292-388.
bool rtransReduce(int distribution, float eta, float alpha, float *extTrans100, float *extDiff, float *intDiff,
                  std::string &err);

// XML subset reader (scenehandler.cpp semantics); fills a HostScene.
bool loadSceneXml(const char *path, const char *const *defines, HostScene &out, std::string &err);

}  // namespace pg
