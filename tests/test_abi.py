"""CPU-side checks of the drop-in boundary: struct layouts, exported symbols, loud failure without a GPU."""
import ctypes as C
import os
import re
import subprocess

import pytest

from dbox_b200 import _abi as A
from dbox_b200 import lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "dbox_b200.h")


def _declared():
    txt = open(HEADER).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(dbx_[a-z0-9_]+)\s*\(", txt)))


def test_struct_sizes_match_header():
    for name, size in A.EXPECTED_SIZES.items():
        assert C.sizeof(getattr(A, name)) == size, name
    # and against the C compiler's view of include/dbox_b200.h
    src = '#include <stdio.h>\n#include "%s"\nint main(){printf("%%zu %%zu %%zu %%zu %%zu %%zu %%zu %%zu", sizeof(dbx_body_def), sizeof(dbx_shape), sizeof(dbx_fixture_def), sizeof(dbx_joint_def), sizeof(dbx_body_state), sizeof(dbx_manifold), sizeof(dbx_contact_rec), sizeof(dbx_proxy_rec));}' % HEADER
    exe = os.path.join(ROOT, ".pytest_cache", "abi_sizes")
    os.makedirs(os.path.dirname(exe), exist_ok=True)
    subprocess.run(["gcc", "-x", "c", "-", "-o", exe], input=src.encode(), check=True)
    got = [int(x) for x in subprocess.check_output([exe]).split()]
    want = [A.EXPECTED_SIZES[k] for k in ("BodyDef", "Shape", "FixtureDef", "JointDef", "BodyState", "Manifold", "ContactRec", "ProxyRec")]
    assert got == want


def test_library_exports_every_declared_symbol():
    lib.build()
    dll = C.CDLL(lib.LIB_PATH)
    names = _declared()
    assert len(names) > 60
    for n in names:
        assert hasattr(dll, n), "libdbox_b200.so does not export " + n
    # the python binding covers the same set
    bound = set("dbx_" + k for k in A.PROTOTYPES)
    assert set(names) == bound, (set(names) ^ bound)


def test_header_cites_reference_for_entry_points():
    txt = open(HEADER).read()
    assert txt.count(".d:") > 40     # file:line citations into /root/reference/src/dbox


def test_no_cpu_fallback_without_device():
    """On a box without a GPU the product must refuse to create a world (and say why), not fall back to anything."""
    api = lib.api()
    if api.device_count() > 0:
        pytest.skip("a CUDA device is visible")
    w = api.world_create(0.0, -10.0, 0, None)
    assert not w
    assert b"no CUDA device" in api.last_error()
    assert api.world_step(None, 1.0 / 60.0, 8, 3) == A.DBX_E_INVALID
    sa = (A.Shape * 1)(); xf = (C.c_float * 4)(0, 0, 0, 1); out = (A.Manifold * 1)()
    api.shape_set_box(C.byref(sa[0]), 1.0, 1.0)
    assert api.debug_collide(0, 1, sa, xf, sa, xf, out) == A.DBX_E_NO_DEVICE


def test_shape_helpers_match_oracle(oracle_api):
    """setup-time host helpers of the shim (SetAsBox / Set / chain) agree bit for bit with the oracle's restatement"""
    import random
    api = lib.api()
    rng = random.Random(7)
    for _ in range(200):
        n = rng.randint(3, 8)
        pts = (A.Vec2 * n)(*[A.Vec2(rng.uniform(-2, 2), rng.uniform(-2, 2)) for _ in range(n)])
        a, b = A.Shape(), A.Shape()
        ca = api.shape_set_polygon(C.byref(a), pts, n)
        cb = oracle_api.shape_set_polygon(C.byref(b), pts, n)
        assert ca == cb
        assert bytes(a)[:200] == bytes(b)[:200]
    a, b = A.Shape(), A.Shape()
    api.shape_set_box_at(C.byref(a), 0.5, 10.0, A.Vec2(10.0, 0.0), 0.3)
    oracle_api.shape_set_box_at(C.byref(b), 0.5, 10.0, A.Vec2(10.0, 0.0), 0.3)
    assert bytes(a)[:200] == bytes(b)[:200]


def test_d_binding_is_generated_from_the_header_and_complete():
    """bindings/d/dbox_b200_c.d (the extern (C) module a D build of dbox imports) is exactly what tools/gen_d_binding.py makes of
    include/dbox_b200.h, declares every entry point once, mirrors every struct field for field, and uses no D keyword as a name"""
    import importlib.util
    spec = importlib.util.spec_from_file_location("gen_d_binding", os.path.join(ROOT, "tools", "gen_d_binding.py"))
    gen = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(gen)
    text, funcs = gen.generate()
    assert open(gen.OUT).read() == text, "run python tools/gen_d_binding.py"
    assert sorted(funcs) == _declared() and len(set(funcs)) == len(funcs)
    for name, typ in (("dbx_body_def", A.BodyDef), ("dbx_shape", A.Shape), ("dbx_joint_def", A.JointDef), ("dbx_contact_rec", A.ContactRec),
                      ("dbx_post_solve", A.PostSolve), ("dbx_world_manifold", A.WorldManifold), ("dbx_ray_hit", A.RayHit)):
        body = re.search(r"struct %s\n\{(.*?)\n\}" % name, text, re.S).group(1)
        assert len([l for l in body.split("\n") if l.strip()]) == len(typ._fields_), name
    for m in re.finditer(r"[\s\*\]]([A-Za-z_]\w*)\s*[;,)]", text.split("extern (C) nothrow @nogc:")[1]):
        assert m.group(1) not in gen.D_KEYWORDS or m.group(1) in ("float", "int", "uint", "long", "ulong", "short", "ushort", "void", "char"), m.group(1)
