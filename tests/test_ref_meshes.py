"""Mesh ingestion against the reference's own loaders: the same .obj / .serialized file goes through
(a) the product's scene reader (xml_scene.cpp, via b200pg_scene_load_xml -> the flat description the device is built from) and
(b) the reference's `obj` plugin (src/shapes/obj.cpp) / TriMesh(Stream *, shapeIndex) (trimesh.cpp:79-270) + TriMesh::configure,
compiled from /root/reference into oracle/_ref. The two may number their vertices differently (the obj loaders de-duplicate
v/vt/vn triples through differently ordered maps), so the comparison is per triangle CORNER: position, shading normal and
texture coordinate of every corner of every triangle, in file order."""
import os
import struct
import zlib

import numpy as np
import pytest

import ref_lib

pytestmark = pytest.mark.skipif(not ref_lib.available(), reason="oracle/_ref not built (no /root/reference here)")

SCENE = """<scene version="0.6.0"><integrator type="progressivepath"/>
<sensor type="perspective"><sampler type="independent"><integer name="sampleCount" value="4"/></sampler>
<film type="hdrfilm"><integer name="width" value="16"/><integer name="height" value="16"/></film></sensor>
%s<shape type="rectangle"><emitter type="area"><rgb name="radiance" value="1"/></emitter></shape></scene>"""


@pytest.fixture(scope="module")
def api(pkg):
    from b200pg import api as _api

    return _api


def corners(P, N, UV, T):
    c = [P[T].reshape(-1, 3)]
    c.append(N[T].reshape(-1, 3) if N is not None else np.zeros((T.size, 3), np.float32))
    c.append(UV[T].reshape(-1, 2) if UV is not None else np.zeros((T.size, 2), np.float32))
    return np.concatenate(c, 1)


def ours(api, tmp_path, shape_xml):
    path = tmp_path / "scene.xml"
    path.write_text(SCENE % shape_xml)
    sc = api.Scene.load_xml(str(path))
    sh = sc.desc.shapes[0]
    nv, nt = sh.n_vertices, sh.n_triangles
    P = np.ctypeslib.as_array(sh.positions, (nv * 3,)).reshape(-1, 3).copy()
    N = np.ctypeslib.as_array(sh.normals, (nv * 3,)).reshape(-1, 3).copy() if bool(sh.normals) else None
    UV = np.ctypeslib.as_array(sh.texcoords, (nv * 2,)).reshape(-1, 2).copy() if bool(sh.texcoords) else None
    T = np.ctypeslib.as_array(sh.indices, (nt * 3,)).reshape(-1, 3).copy()
    sc.close()
    return corners(P, N, UV, T)


def theirs(kind, path, **kw):
    ms = ref_lib.load_meshes(kind, path, **kw)
    return np.concatenate([corners(m["P"], m["N"], m["UV"], m["T"]) for m in ms], 0)


OBJ = """# a bent strip: shared vertices, a quad, explicit normals and texture coordinates on some faces
v 0 0 0
v 1 0 0
v 1 1 0.2
v 0 1 0.1
v 2 0 0.5
v 2 1 0.7
vt 0 0
vt 1 0
vt 1 1
vt 0 1
vn 0 0 1
vn 0.1 0 0.99
f 1/1/1 2/2/1 3/3/2 4/4/2
f 2/2/1 5/1/2 6/4/2
f 2/2/1 6/4/2 3/3/2
"""

OBJ_PLAIN = """v 0 0 0
v 1 0 0
v 1 1 0.2
v 0 1 0.1
v 2 0 0.5
v 2 1 0.7
f 1 2 3
f 1 3 4
f 2 5 6
f -5 -1 -4
"""


@pytest.mark.parametrize("text,opts", [(OBJ, ""), (OBJ_PLAIN, ""), (OBJ_PLAIN, "face"), (OBJ_PLAIN, "flip"), (OBJ, "xform")])
def test_obj_loader_matches_the_reference_plugin(api, tmp_path, text, opts):
    f = tmp_path / "m.obj"
    f.write_text(text)
    extra, kw = "", {}
    if opts == "face":
        extra, kw = '<boolean name="faceNormals" value="true"/>', dict(face_normals=True)
    if opts == "flip":
        extra, kw = '<boolean name="flipNormals" value="true"/>', dict(flip_normals=True)
    if opts == "xform":
        extra = '<transform name="toWorld"><scale x="2" y="0.5" z="3"/><rotate y="1" angle="30"/><translate x="1" y="2" z="-3"/></transform>'
        c, s = np.cos(np.radians(30.0)), np.sin(np.radians(30.0))
        R = np.array([[c, 0, s, 0], [0, 1, 0, 0], [-s, 0, c, 0], [0, 0, 0, 1]])
        M = np.array([[1, 0, 0, 1], [0, 1, 0, 2], [0, 0, 1, -3], [0, 0, 0, 1.0]]) @ R @ np.diag([2, 0.5, 3, 1.0])
        kw = dict(to_world=M)
    a = ours(api, tmp_path, '<shape type="obj"><string name="filename" value="%s"/>%s<bsdf type="diffuse"/></shape>' % (f, extra))
    b = theirs("obj", f, **kw)
    assert a.shape == b.shape
    np.testing.assert_allclose(a, b, atol=3e-6)


def _serialized(path, meshes, version=4):
    """trimesh.cpp:175-270: per shape u16 0x041C, u16 version, zlib stream {u32 flags, [v4: name\\0], u64 nv, u64 nt, positions,
    [normals], [texcoords], indices}; then the offset dictionary of serialized.cpp (u64 offsets for v4, u32 for v3; u32 count)."""
    blob, offs = b"", []
    for (P, T, N, UV) in meshes:
        offs.append(len(blob))
        flags = 0x1000 | (0x0001 if N is not None else 0) | (0x0002 if UV is not None else 0)
        body = struct.pack("<I", flags) + (b"mesh\0" if version == 4 else b"") + struct.pack("<QQ", len(P), len(T))
        body += P.astype("<f4").tobytes()
        if N is not None:
            body += N.astype("<f4").tobytes()
        if UV is not None:
            body += UV.astype("<f4").tobytes()
        body += T.astype("<u4").tobytes()
        blob += struct.pack("<HH", 0x041C, version) + zlib.compress(body)
    for o in offs:
        blob += struct.pack("<Q" if version == 4 else "<I", o)
    blob += struct.pack("<I", len(meshes))
    with open(path, "wb") as f:
        f.write(blob)


@pytest.mark.parametrize("version", [3, 4])
def test_serialized_loader_matches_the_reference(api, pkg, tmp_path, version):
    S = pkg.scenes
    P0, N0, T0 = S.heightfield_mesh(n=7, seed=3, amp=0.2)
    P1, _, T1 = S.heightfield_mesh(n=5, seed=9, amp=0.4)
    UV0 = (P0[:, :2] * 0.5 + 0.25).astype(np.float32)
    f = str(tmp_path / "two.serialized")
    _serialized(f, [(P0.astype(np.float32), T0, N0.astype(np.float32), UV0), (P1.astype(np.float32), T1, None, None)], version)
    for idx in (0, 1):
        a = ours(api, tmp_path, '<shape type="serialized"><string name="filename" value="%s"/><integer name="shapeIndex" value="%d"/>'
                                '<bsdf type="diffuse"/></shape>' % (f, idx))
        b = theirs("serialized", f, shape_index=idx)
        assert a.shape == b.shape
        np.testing.assert_allclose(a, b, atol=3e-6)  # shape 1 has no normals: both sides generate the angle-weighted ones
