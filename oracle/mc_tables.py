"""ORACLE / build-time generator — test infrastructure and table source, NOT product code at run time.

Marching-cubes case table, GENERATED (not transcribed): for each of the 256 corner-sign configurations of a cell the isosurface
polygons are found by tracing, on each of the 6 cube faces, the segments that separate inside (value < level) from outside
corners, chaining the segments into closed loops over the 12 cube edges and fanning each loop into triangles.  A face with two
diagonally opposite inside corners (the ambiguous face) always cuts each inside corner off separately -- a rule that depends on
the face's four corners only, so the two cells sharing the face agree and the mesh is watertight.

The reference calls skimage.measure.marching_cubes_lewiner (model/sdf_net.py:103), a third-party routine that is absent from this
image (and removed from current scikit-image); its exact triangulation is unpinned here.  What is pinned (tests/test_mc_*.py): the
extracted surface is closed and consistently oriented, every vertex lies on a cell edge at the linearly interpolated level crossing,
and GPU output == this module's numpy implementation, vertex for vertex and face for face.

Conventions: corner c = x + 2 y + 4 z (bits); edge e = axis * 4 + (a + 2 b) with (a, b) the corner's coordinates on the other two
axes in ascending axis order; triangles are oriented so that their normal points towards increasing value (out of the shape for an SDF)."""
import numpy as np

AXES = ((1, 0, 0), (0, 1, 0), (0, 0, 1))


def corner(x, y, z):
    return x + 2 * y + 4 * z


def edge_id(axis, a, b):
    return axis * 4 + a + 2 * b


def edge_corners(e):
    axis, k = divmod(e, 4)
    a, b = k & 1, k >> 1
    p = [0, 0, 0]
    others = [i for i in range(3) if i != axis]
    p[others[0]], p[others[1]] = a, b
    q = list(p)
    q[axis] = 1
    return corner(*p), corner(*q)


def _edge_between(c0, c1):
    d = c0 ^ c1
    axis = {1: 0, 2: 1, 4: 2}[d]
    lo = min(c0, c1)
    p = [(lo >> i) & 1 for i in range(3)]
    others = [i for i in range(3) if i != axis]
    return edge_id(axis, p[others[0]], p[others[1]])


def _faces():
    """each face: 4 corners in counter-clockwise order when seen from OUTSIDE the cube"""
    faces = []
    for axis in range(3):
        u, v = [i for i in range(3) if i != axis]
        for side in (0, 1):
            cyc = []
            for (a, b) in ((0, 0), (1, 0), (1, 1), (0, 1)):
                p = [0, 0, 0]
                p[axis], p[u], p[v] = side, a, b
                cyc.append(corner(*p))
            # (u, v, axis) is a cyclic permutation of (x, y, z) for axis = 2, 0; for axis = 1 (u, v) = (x, z) is odd.
            # seen from +axis the order (0,0),(1,0),(1,1),(0,1) in (u, v) is CCW iff (u, v, axis) is right-handed
            right_handed = (u, v, axis) in ((0, 1, 2), (1, 2, 0), (2, 0, 1))
            ccw_from_plus = right_handed
            if (side == 1) != ccw_from_plus:
                cyc.reverse()
            faces.append(cyc)
    return faces


FACES = _faces()


def _segments(case):
    """directed segments (edge_from, edge_to); direction: walking along the segment seen from outside the cube, INSIDE is on the left"""
    segs = []
    inside = [(case >> c) & 1 for c in range(8)]
    for cyc in FACES:
        ins = [inside[c] for c in cyc]
        n = sum(ins)
        if n == 0 or n == 4:
            continue
        # edges of the face between consecutive corners; crossing where the signs differ
        def e(i):
            return _edge_between(cyc[i], cyc[(i + 1) % 4])
        if n == 2 and ins[0] == ins[2]:
            # ambiguous face: cut every inside corner off on its own
            for i in range(4):
                if ins[i]:
                    segs.append((e((i - 1) % 4), e(i)))
            continue
        # one contiguous run of inside corners i0 .. i1 (cyclic): the segment enters through the edge before the run and leaves
        # through the edge after it; with CCW corners seen from outside, going from edge(i0-1) to edge(i1) keeps the inside run on the left
        start = [i for i in range(4) if ins[i] and not ins[(i - 1) % 4]][0]
        end = start
        while ins[(end + 1) % 4]:
            end = (end + 1) % 4
        segs.append((e((start - 1) % 4), e(end)))
    return segs


def build_table():
    """tri[case] = list of (e0, e1, e2) edge triples"""
    table = []
    for case in range(256):
        segs = _segments(case)
        nxt = {}
        for a, b in segs:
            assert a not in nxt, (case, segs)
            nxt[a] = b
        tris = []
        seen = set()
        for start in sorted(nxt):
            if start in seen:
                continue
            loop = [start]
            seen.add(start)
            cur = nxt[start]
            while cur != start:
                loop.append(cur)
                seen.add(cur)
                cur = nxt[cur]
            assert len(loop) >= 3, (case, loop)
            tris.extend(_triangulate(loop))
        table.append(tris)
    return table


def _edge_faces(e):
    c0, c1 = edge_corners(e)
    return {i for i, cyc in enumerate(FACES) if c0 in cyc and c1 in cyc}


def _triangulate(loop):
    """Triangles of a closed loop of cube-edge vertices.  A diagonal joining two vertices of the same cube face would lie IN that face,
    where the neighbouring cell can produce the mirrored triangle (a zero-volume fin): prefer triangulations without such diagonals."""
    n = len(loop)
    cands = []
    for r in range(n):                                   # fans from every apex
        q = loop[r:] + loop[:r]
        cands.append([(q[0], q[i], q[i + 1]) for i in range(1, n - 1)])
    if n == 6:                                           # inner triangle + three ears
        for r in (0, 1):
            q = loop[r:] + loop[:r]
            cands.append([(q[0], q[2], q[4]), (q[0], q[1], q[2]), (q[2], q[3], q[4]), (q[4], q[5], q[0])])

    def in_face_diagonals(tris):
        adj = {(loop[i], loop[(i + 1) % n]) for i in range(n)}
        bad = 0
        for t in tris:
            for i in range(3):
                a, b = t[i], t[(i + 1) % 3]
                if (a, b) in adj or (b, a) in adj:
                    continue
                if _edge_faces(a) & _edge_faces(b):
                    bad += 1
        return bad
    return min(cands, key=in_face_diagonals)


def _orientation_sign(table):
    """+1 if the triangles of the single-inside-corner case face away from the inside corner (towards increasing value)"""
    tri = table[1][0]                      # corner 0 inside
    mid = []
    for e in tri:
        c0, c1 = edge_corners(e)
        p0 = np.array([(c0 >> i) & 1 for i in range(3)], dtype=float)
        p1 = np.array([(c1 >> i) & 1 for i in range(3)], dtype=float)
        mid.append((p0 + p1) / 2)
    n = np.cross(mid[1] - mid[0], mid[2] - mid[0])
    return 1 if np.dot(n, np.array([1.0, 1.0, 1.0])) > 0 else -1


def tables():
    """(tri_count uint8 [256], tri_edges int8 [256][MAX*3] padded with -1, edge_corner uint8 [12][2])"""
    t = build_table()
    if _orientation_sign(t) < 0:
        t = [[(a, c, b) for (a, b, c) in tris] for tris in t]
    mx = max(len(x) for x in t)
    count = np.array([len(x) for x in t], dtype=np.uint8)
    edges = -np.ones((256, mx * 3), dtype=np.int8)
    for i, tris in enumerate(t):
        for j, tri in enumerate(tris):
            edges[i, 3 * j:3 * j + 3] = tri
    ec = np.array([edge_corners(e) for e in range(12)], dtype=np.uint8)
    return count, edges, ec


def marching_cubes(volume, level=0.0, spacing=(1.0, 1.0, 1.0)):
    """numpy reference of the GPU pipeline (same tables, same fp32 interpolation, same ordering):
    vertices are owned by the cell at the lower end of their edge, numbered in cell order then axis order; faces in cell order.
    Returns (vertices float32 [V,3] in index space * spacing, faces int32 [F,3], normals float32 [V,3])."""
    vol = np.ascontiguousarray(volume, dtype=np.float32)
    count, tri_edges, ec = tables()
    nx, ny, nz = vol.shape
    lvl = np.float32(level)
    inside = vol < lvl
    # ---- vertices: edge (cell p, axis a) crosses iff inside differs between p and p + e_a
    vid = -np.ones((nx, ny, nz, 3), dtype=np.int64)
    cross = np.zeros((nx, ny, nz, 3), dtype=bool)
    cross[:-1, :, :, 0] = inside[:-1] != inside[1:]
    cross[:, :-1, :, 1] = inside[:, :-1] != inside[:, 1:]
    cross[:, :, :-1, 2] = inside[:, :, :-1] != inside[:, :, 1:]
    flat = cross.reshape(-1)
    ids = np.cumsum(flat) - 1
    vid.reshape(-1)[flat] = ids[flat]
    cells = np.argwhere(cross)                       # rows (x, y, z, axis) in C order == numbering order
    p = cells[:, :3]
    a = cells[:, 3]
    q = p.copy()
    q[np.arange(len(q)), a] += 1
    v0 = vol[p[:, 0], p[:, 1], p[:, 2]]
    v1 = vol[q[:, 0], q[:, 1], q[:, 2]]
    t = ((lvl - v0) / (v1 - v0)).astype(np.float32)
    verts = p.astype(np.float32)
    verts[np.arange(len(verts)), a] += t
    sp = np.asarray(spacing, dtype=np.float32)
    # ---- normals: central-difference gradient at both ends of the edge, interpolated, normalised
    g = np.stack(np.gradient(vol.astype(np.float32)), axis=-1).astype(np.float32) if min(vol.shape) > 1 else np.zeros(vol.shape + (3,), np.float32)
    g0 = g[p[:, 0], p[:, 1], p[:, 2]]
    g1 = g[q[:, 0], q[:, 1], q[:, 2]]
    nrm = (g0 + (g1 - g0) * t[:, None]) / sp
    ln = np.linalg.norm(nrm, axis=1, keepdims=True)
    nrm = np.where(ln > 0, nrm / np.maximum(ln, 1e-30), 0).astype(np.float32)
    # ---- faces
    cx, cy, cz = nx - 1, ny - 1, nz - 1
    case = np.zeros((cx, cy, cz), dtype=np.int32)
    for c in range(8):
        dx, dy, dz = c & 1, (c >> 1) & 1, (c >> 2) & 1
        case |= inside[dx:dx + cx, dy:dy + cy, dz:dz + cz].astype(np.int32) << c
    faces = []
    act = np.argwhere(count[case] > 0)
    for (x, y, z) in act:
        k = case[x, y, z]
        for j in range(count[k]):
            tri = []
            for e in tri_edges[k, 3 * j:3 * j + 3]:
                axis, kk = divmod(int(e), 4)
                aa, bb = kk & 1, kk >> 1
                o = [x, y, z]
                others = [i for i in range(3) if i != axis]
                o[others[0]] += aa
                o[others[1]] += bb
                tri.append(vid[o[0], o[1], o[2], axis])
            faces.append(tri)
    faces = np.array(faces, dtype=np.int32).reshape(-1, 3)
    return verts * sp, faces, nrm
