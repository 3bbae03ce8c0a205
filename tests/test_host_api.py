"""Host-side logic of the product package (CPU only, no compute calls)."""
import ctypes
import os
import re

import numpy as np
import pytest

from conftest import ROOT, load_golden
from helpers import prisms_from_golden

from fatiando_b200 import constants, gridder, mesher, utils
from fatiando_b200.flatten import flatten


def test_support_functions_match_reference():
    d = load_golden("support")
    for (inc, dec), want in zip(d['angles'], d['dircos']):
        assert np.array_equal(np.array(utils.dircos(inc, dec)), want)
    for (inc, dec), want in zip(d['angles'], d['ang2vec_2p5']):
        assert np.array_equal(utils.ang2vec(2.5, inc, dec), want)
    assert np.array_equal(utils.ang2vec(np.arange(4.), 45, 45), d['ang2vec_array'])
    x, y, z = gridder.regular((0, 10, -3, 5), (5, 3), z=-10)
    assert np.array_equal(x, d['regular_x']) and np.array_equal(y, d['regular_y'])
    assert np.array_equal(z, d['regular_z'])
    assert gridder.spacing((0, 10, 0, 20), (5, 21)) == list(d['spacing'])
    C = constants
    got = np.array([C.G, C.SI2MGAL, C.SI2EOTVOS, C.CM, C.T2NT, C.G * C.SI2MGAL,
                    C.G * C.SI2EOTVOS, C.CM * C.T2NT])
    assert np.array_equal(got, d['constants'])


def _mesh_from(d):
    m = mesher.PrismMesh(tuple(d['mesh_bounds']), tuple(int(v) for v in d['mesh_shape']))
    return m


def test_prism_mesh_iteration_and_flatten_bitwise():
    d = load_golden("prism_mesh")
    m = _mesh_from(d)
    m.addprop('density', d['mesh_density'])
    m.mask = [int(v) for v in d['mesh_mask']]
    assert tuple(m.dims) == tuple(d['mesh_dims'])
    # object iteration == reference iteration
    cells = list(m)
    assert [c is None for c in cells] == list(d['isnone'])
    for c, row, dens in zip(cells, d['bounds'], d['density']):
        if c is not None:
            assert c.get_bounds() == list(row) and c.props['density'] == dens
    # vectorised flattening == reference iteration, bit for bit
    flat = flatten(m, 'density')
    keep = ~d['isnone']
    assert np.array_equal(flat.index, np.nonzero(keep)[0]) and flat.total == m.size
    for k in range(6):
        assert np.array_equal(flat.bounds[k], d['bounds'][keep][:, k])
    assert np.array_equal(flat.density, d['density'][keep])
    # bounds that are not representable: x2 = x1 + dx, never b0 + dx*(i+1)
    m2 = mesher.PrismMesh(tuple(d['awkward_mesh_bounds']), tuple(int(v) for v in d['awkward_mesh_shape']))
    flat2 = flatten(m2, None, override=True)
    for k in range(6):
        assert np.array_equal(flat2.bounds[k], d['awkward_bounds'][:, k])
    # a mesh without the property is skipped entirely unless overridden
    assert flatten(m2, 'density').size == 0
    assert flatten(m2, 'density', override=True).size == m2.size


def test_prism_mesh_magnetization_flatten():
    d = load_golden("mesh_magnetic")
    m = _mesh_from(d)
    m.addprop('magnetization', [row for row in d['mesh_magnetization']])
    flat = flatten(m, 'magnetization')
    assert np.array_equal(flat.magnetization, d['magnetization'])
    for k in range(6):
        assert np.array_equal(flat.bounds[k], d['bounds'][:, k])
    # scalar (induced) magnetization is expanded along F only when F is known
    m.addprop('magnetization', [2.0] * m.size)
    fv = utils.dircos(30, -15)
    flat = flatten(m, 'magnetization', fvec=fv)
    assert np.array_equal(flat.magnetization[0], np.array([2.0 * fv[0], 2.0 * fv[1], 2.0 * fv[2]]))
    with pytest.raises(TypeError):
        flatten(m, 'magnetization')


def test_prism_relief_flatten_bitwise():
    d = load_golden("prism_relief")
    rel = mesher.PrismRelief(float(d['relief_ref']), tuple(d['relief_dims']),
                             (d['relief_x'], d['relief_y'], d['relief_z']))
    rel.addprop('density', d['relief_density_in'])
    flat = flatten(rel, 'density')
    for k in range(6):
        assert np.array_equal(flat.bounds[k], d['bounds'][:, k])
    assert np.array_equal(flat.density, d['density'])       # sign flipped where z > ref
    assert (flat.density < 0).any() and (flat.density > 0).any()
    p = rel[3]
    assert p.get_bounds() == list(d['bounds'][3]) and p.props['density'] == d['density'][3]


def test_generic_list_flatten_and_skipping():
    d = load_golden("skip_override")
    model = prisms_from_golden(d, mesher.Prism)
    assert model[0] is None and model[3] is None
    flat = flatten(model, 'density')
    assert list(flat.index) == [1, 4] and flat.total == 5
    flat = flatten(model, 'magnetization')
    assert list(flat.index) == [1, 2] and flat.magnetization.shape == (2, 3)
    flat = flatten(model, 'density', override=True)
    assert list(flat.index) == [1, 2, 4] and flat.density is None
    assert flatten([], 'density').size == 0
    assert flatten([None, None], 'density').size == 0


def test_mesh_conveniences():
    m = mesher.PrismMesh((0, 2, 0, 4, 0, 3), (1, 1, 2))
    assert list(m.get_xs()) == [0., 1., 2.] and list(m.get_ys()) == [0., 4.]
    assert str(m[-1]) == 'x1:1 | x2:2 | y1:0 | y2:4 | z1:0 | z2:3'
    with pytest.raises(IndexError):
        m[2]
    with pytest.raises(AttributeError):
        mesher.PrismMesh((0, 2, 0, 4, 0, 3), (1, 1, 2.5))
    m = mesher.PrismMesh((0, 2, 0, 2, 0, 2), (2, 2, 2))
    assert [str(p) for p in m.get_layer(1)][0] == 'x1:0 | x2:1 | y1:0 | y2:1 | z1:1 | z2:2'
    assert len(list(m.layers())) == 2
    p = mesher.Prism(1, 2, 1, 3, 0, 2)
    assert list(p.center()) == [1.5, 2.0, 1.0]


def test_shape_mismatch_raises_before_touching_the_gpu():
    """test_prism.py:11-118 of the reference, re-pointed at the new module."""
    from fatiando_b200 import prism
    inc, dec = 10, 0
    model = [mesher.Prism(-6000, -2000, 2000, 4000, 0, 3000,
                          {'density': 1000, 'magnetization': utils.ang2vec(10, inc, dec)})]
    x, y, z = gridder.regular([-5000, 5000, -10000, 10000], (101, 51), z=-1)
    bad = [(x[:-2], y, z), (x, y[:-2], z), (x, y, z[:-2]), (x[:-5], y, z[:-2])]
    for name in prism.GRAVITY_FIELDS + ('bx', 'by', 'bz'):
        for args in bad:
            with pytest.raises(ValueError):
                getattr(prism, name)(*args, model)
    for args in bad:
        with pytest.raises(ValueError):
            prism.tf(*args, model, inc, dec)
        for k in ('xx', 'xy', 'xz', 'yy', 'yz', 'zz'):
            with pytest.raises(ValueError):
                getattr(prism, 'kernel' + k)(*args, model[0])


def test_library_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, "include", "fatiando_b200.h")).read()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    declared = set(re.findall(r"\b(fb200_\w+)\s*\(", header))
    assert len(declared) >= 28
    from fatiando_b200 import _lib
    L = ctypes.CDLL(_lib.LIB_PATH)      # loading needs no GPU
    missing = [s for s in sorted(declared) if not hasattr(L, s)]
    assert not missing, missing
    assert set(_lib._SIGNATURES) == declared
    assert _lib.lib().fb200_abi_version() == 1


def test_no_cpu_fallback_without_a_device():
    """On a machine without a GPU the product must fail loudly, not compute."""
    from fatiando_b200 import _lib, prism
    if _lib.device_count() > 0:
        pytest.skip("a CUDA device is present")
    model = [mesher.Prism(0, 1, 0, 1, 0, 1, {'density': 1.0})]
    x = np.zeros(3); y = np.zeros(3); z = -np.ones(3)
    with pytest.raises(RuntimeError):
        prism.gz(x, y, z, model)
    assert "CUDA" in _lib.last_error() or "device" in _lib.last_error()


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "fatiando_b200")
    for dirpath, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, fn)).read()
                assert "oracle" not in text.lower().replace("not the oracle", ""), fn


def test_doctests_of_the_host_modules():
    import doctest
    from fatiando_b200 import utils, mesher
    for mod in (utils, mesher):
        res = doctest.testmod(mod)
        assert res.failed == 0 and res.attempted > 0, mod.__name__


def test_kernel_profiles_are_valid_for_this_tree():
    """
    bench.py's roofline reads executed-instruction counts from profiles/r02_kernel_<key>.json and
    refuses a file measured on other DEVICE sources of that kernel (bench.load_profile).  Every
    committed profile must be valid for the committed sources: a kernel change without a
    re-capture (tools/profile_all.py) fails here, on the CPU, before a bench line carries
    `frac: null`.  The profiles the default bench line quotes must also have been captured with
    the committed LAUNCH code (grid sizing, prism split).
    """
    import glob
    import json
    import sys
    sys.path.insert(0, ROOT)
    import bench
    files = sorted(glob.glob(os.path.join(ROOT, "profiles", "r02_kernel_*.json")))
    assert len(files) >= 14
    for path in files:
        d = json.load(open(path))
        key = d["key"]
        assert key in bench.KERNEL_TU, key
        names = [os.path.basename(f) for f in bench.kernel_sources(key)]
        assert bench.KERNEL_TU[key] in names and "prism_common.cuh" in names
        prof, src = bench.load_profile(key)
        assert prof is not None, src
    for key in ("c2_nodes", "c3_nodes", "c4_nodes", "c4_grid", "c5_surface", "c5t_surface"):
        prof, src = bench.load_profile(key)
        assert prof["launch_code"] == "as captured", (key, prof["launch_code"])
    assert [os.path.basename(f) for f in bench.launch_sources()] == ["prism_abi.cu", "abi_internal.h", "fatiando_b200.h"]
    # a header of another kernel family is not part of a prism kernel's hash
    assert "polyprism.cuh" not in [os.path.basename(f) for f in bench.kernel_sources("c2_nodes")]
    assert "polyprism.cuh" in [os.path.basename(f) for f in bench.kernel_sources("polyprism_gz")]


def test_point_grid_and_mesh_dump_match_the_reference():
    """PointGrid (mesh.py:186-388) and PrismMesh.dump (mesh.py:831-891) against the real reference's outputs."""
    import io
    g = dict(np.load(os.path.join(ROOT, "tests", "golden", "containers.npz")))
    grid = mesher.PointGrid(tuple(g['pg_area']), g['pg_zin'], tuple(int(v) for v in g['pg_shape']))
    grid.addprop('density', g['pg_density'])
    for name, ours in (('pg_x', grid.x), ('pg_y', grid.y), ('pg_z', grid.z)):
        assert np.array_equal(g[name], ours), name
    assert grid.radius == float(g['pg_radius']) and [grid.dx, grid.dy] == list(g['pg_spacing'])
    assert len(grid) == 24 and len(list(grid)) == 24
    items = [grid[i] for i in (0, 5, -1)]
    assert np.array_equal(g['pg_items'], [[s.x, s.y, s.z, s.radius, s.props['density']] for s in items])
    parts = grid.split((3, 2))
    assert len(parts) == int(g['pg_split_n'])
    for k, s in enumerate(parts):
        assert np.array_equal(g['pg_split_%d' % k], [s.x, s.y, s.z, s.props['density']]), k
        assert np.array_equal(g['pg_split_area_%d' % k], s.area), k
    with pytest.raises(ValueError):
        grid.split((4, 2))
    with pytest.raises(IndexError):
        grid[24]
    flat = mesher.PointGrid((0, 10, 2, 6), 200, (2, 3))
    assert np.array_equal(g['pg_flat'], [flat.x, flat.y, flat.z])
    # the sphere flattener takes the grid's arrays as they are
    from fatiando_b200 import sphere
    geo, dens, mag = sphere._flatten(grid, 'density', False)
    listed = sphere._flatten(list(grid), 'density', False)
    assert all(np.array_equal(a, b) for a, b in zip(geo, listed[0])) and np.array_equal(dens, listed[1])

    mesh = mesher.PrismMesh(tuple(g['dump_bounds']), tuple(int(v) for v in g['dump_shape']))
    mesh.addprop('density', g['dump_density'])
    for tag, mask in (('plain', []), ('masked', g['dump_mask'].tolist())):
        mesh.mask = list(mask)
        mf, pf = io.StringIO(), io.StringIO()
        mesh.dump(mf, pf, 'density')
        assert mf.getvalue() == g['dump_%s_mesh' % tag].tobytes().decode(), tag
        assert pf.getvalue() == g['dump_%s_prop' % tag].tobytes().decode(), tag
    with pytest.raises(ValueError):
        mesh.dump(io.StringIO(), io.StringIO(), 'magnetization')
