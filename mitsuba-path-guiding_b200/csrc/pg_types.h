// pg_types.h -- POD records shared by the host scene compiler and the CUDA kernels.
// Layouts are chosen for 16-byte vector loads (LDG.128): every record is a whole number of float4s.
#pragma once
#include <stdint.h>

#include "../../include/b200pg.h"

namespace pg {

static const uint32_t kNoTriangle = 0xFFFFFFFFu;  // rectangle marker, as KNoTriangleFlag in skdtree.h
static const uint32_t kMiss = 0xFFFFFFFFu;

// BVH2 node, 64 B = 4 x float4 (one 128-B line holds two nodes).
//   q0 = c0.min.x c0.max.x c0.min.y c0.max.y
//   q1 = c1.min.x c1.max.x c1.min.y c1.max.y
//   q2 = c0.min.z c0.max.z c1.min.z c1.max.z
//   q3 = child0, child1 (int bits), pad, pad.   child >= 0: inner node index;
//        child < 0: leaf, ~child = (first_prim << 7) | (rectMask << 3) | count   (count <= 4;
//        bit i of rectMask: primitive first_prim + i is a rectangle)
struct BvhNode {
    float q[16];
};
static const int kLeafShift = 7;

// Wide BVH node, 96 B = 6 x float4 = three 32-byte sectors: up to 8 children, boxes quantised to 8 bits on the node's own grid.
//   q0     = origin.xyz (float), meta = ex | ey << 8 | ez << 16 | nChildren << 24   (e*: biased exponent of the grid step 2^e)
//   q1, q2 = child[8] (int bits): >= 0 wide node index, < 0 leaf code (as in BvhNode), 0x7FFFFFFF empty slot
//   q3..q5 = bytes lox[8] loy[8] loz[8] hix[8] hiy[8] hiz[8]: plane position = origin + q * 2^e
// Built by collapsing the binary tree (host_scene.cpp); conservative: every child box is widened by >= 1 grid step.
struct WideNode {
    float q[24];
};
static const int kWideMaxDepth = 16;   // the traversal stack holds 7 * depth + postponed entries (device_scene.cuh: kWideStack)
static const int32_t kWideEmpty = 0x7FFFFFFF;

// Primitive record, 48 B = 3 x float4 (rows of a 3x4 affine map), stored in BVH leaf order.
// One branch-free test serves both primitive kinds: with l(x) = M x + w,
//     t = -l(o).z / (M d).z,   (u, v) = (l(o) + t M d).xy
//   rectangle: M, w = worldToObject (the arithmetic of rectangle.cpp:125-148); accept |u|,|v| <= 1
//   triangle:  row2 = geometric plane, rows 0/1 = barycentric planes of vertices 1 and 2;
//              accept u, v >= 0, u + v <= 1 (same (t,u,v) as TriAccel::rayIntersect, triaccel.h:96-158,
//              up to rounding)
struct PrimRecord {
    float q[12];
};

// Shading record of a primitive, 96 B = 6 x float4 = three whole 32-byte sectors, in BVH leaf order (indexed by the hit's
// primitive slot). Everything fillIntersection needs about a triangle sits in ONE place: the previous chain
// PrimInfo -> ShapeRecord -> MeshRecord -> indices -> positions / normals was five dependent, scattered loads deep and
// touched ~9 sectors per hit on the 10 M-triangle mesh.
//   r0 = p0.xyz, shape index       r1 = p1.xyz, flags (bit 0: triangle, bit 1: has vertex normals, bit 2: has a UV tangent,
//                                                      bits 8..31: emitter index + 1, 0 = none)
//   r2 = p2.xyz, primitive index   r3 = n0.xyz, bsdf index   r4 = n1.xyz, a   r5 = n2.xyz, b
// (a, b): the triangle's UV tangent dpdu = a (p1 - p0) + b (p2 - p0) of TriMesh::computeUVTangents (trimesh.cpp:683-735), which
// fillIntersectionRecord uses instead of the first edge whenever the mesh has texture coordinates (skdtree.h:374-381).
// Rectangles keep only the .w words of r0..r3 (their frame lives in RectRecord).
struct ShadeTri {
    float q[24];
};

struct PrimInfo {  // (shape, primitive) of a BVH slot: medium-boundary logic of the volumetric path
    uint32_t shape, prim;  // prim == kNoTriangle for rectangles
};

// Rectangle, 8 x float4: rows of worldToObject (3), then frame/dpdu data for shading.
//   r0..r2 = worldToObject rows (x y z w)
//   r3 = n.xyz, invSurfaceArea   r4 = dpdu.xyz, 0   r5 = objectToWorld row0   r6 = row1   r7 = row2
struct RectRecord {
    float q[32];
};

struct ShapeRecord {  // 32 B
    int32_t type;      // B200pgShapeType
    int32_t bsdf;
    int32_t emitter;
    int32_t interiorMedium, exteriorMedium;
    uint32_t primOffset;   // first global primitive id
    uint32_t meshOffset;   // trimesh: index into the mesh table; rectangle: rect index
    uint32_t flags;        // bit0 = has vertex normals
};

struct MeshRecord {  // offsets into the concatenated vertex/index pools
    uint32_t vertexOffset;  // in vertices
    uint32_t indexOffset;   // in triangles
    uint32_t cdfOffset;     // area cdf (n_triangles + 1 floats)
    uint32_t nTriangles;
    float invSurfaceArea;
    uint32_t hasNormals;
    uint32_t hasTexcoords;
    uint32_t pad1;
};

struct BsdfRecord {  // mirrors B200pgBsdf + derived constants
    int32_t type, twosided;
    float reflectance[3];
    float specRefl[3];
    float specTrans[3];
    float eta, invEta;          // intIOR / extIOR
    float condEta[3], condK[3];
    int32_t distribution;
    float alphaU, alphaV;
    int32_t nonlinear;
    float specSamplingWeight;   // roughplastic.cpp:283-286
    float invEta2;
    float rtIntDiff;
    uint32_t typeFlags;         // BSDF::getType() bits
    float rtExt[100];           // reduced rough transmittance (rtrans.h:292-388)
    float pad[3];
};

struct EmitterRecord {
    float radiance[3];
    int32_t shape;
};

struct CameraRecord {
    float sampleToCamera[16];
    float toWorld[12];  // 3x4 affine
    float nearClip, farClip;
    float invResX, invResY;
    int32_t medium;
};

struct FilmRecord {
    int32_t width, height;
    float radius, scaleFactor;
    float values[32];
};

struct MediumRecord {
    int32_t method, phaseType;
    float scale, invMaxDensity, maxDensity, g;
    float albedo[3];
    int32_t res[3];
    float aabbMin[3], aabbMax[3];
    float worldToGrid[12];  // 3x4 affine
    uint64_t densityOffset; // into the density pool (floats)
    float stepSize;
    float pad;
};

}  // namespace pg
