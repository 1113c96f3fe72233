"""The MJCF -> structure-of-arrays compiler (mjcf/parser.py + mjcf/compile.py) run FROM THE XML: every other test loads the
committed models/*.npz, so a regression in the parser or the compiler would go unnoticed.  Needs the
reference checkout (the MJCF / STL assets are the reference's own and are not copied into this repo): skipped on the GPU
box, run in the build container."""
import os
import sys

import numpy as np
import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("EARL_REFERENCE", "/root/reference")
pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "earl_benchmark", "envs", "metaworld_assets")),
                                reason="needs the reference checkout for the MJCF / STL assets")


@pytest.fixture(scope="module")
def CM():
    sys.path.insert(0, os.path.join(REPO, "tools"))
    import compile_models
    return compile_models


@pytest.mark.parametrize("name", ["sawyer_door", "sawyer_peg", "kitchen"])
def test_compile_from_xml_reproduces_the_committed_model(CM, name):
    from earl_benchmark_b200.mjcf.compile import Model
    fresh = getattr(CM, name)()
    fresh._ext_defaults()
    stored = Model.load(os.path.join(REPO, "earl_benchmark_b200", "models", name + ".npz"))
    stored._ext_defaults()
    for k, _ in Model.FIELDS + Model.EXT_FIELDS:
        a, b = np.asarray(getattr(fresh, k)), np.asarray(getattr(stored, k))
        assert a.shape == b.shape, k
        assert np.array_equal(a, b), (name, k, np.abs(a.astype(float) - b.astype(float)).max())
    assert fresh.names == stored.names
    assert fresh.to_blob() == stored.to_blob()


def test_parser_semantics_on_the_sawyer_scene(CM):
    """Spot checks of what the parser must get right for the physics to be MuJoCo's: nested default
    classes through `childclass`, last-one-wins <compiler>, explicit <inertial> against inertiafromgeom="auto" with
    inertiagrouprange 4-5, mocap bodies, weld equality, position actuators."""
    from earl_benchmark_b200.mjcf import compile as C, parser
    spec = parser.load(os.path.join(CM.MW, "sawyer_door_pull.xml"))
    assert spec.compiler["angle"] == "radian" and spec.compiler["inertiagrouprange"] == "4 5"
    assert spec.option["timestep"] == "0.0025" and spec.option["cone"] == "elliptic" and spec.option["iterations"] == "50"
    by = {b["name"]: b for b in spec.bodies}
    pad = {g.get("name"): g for g in by["rightpad"]["geoms"]}["rightpad_geom"]
    assert pad["group"] == "1" and pad["condim"] == "4" and pad["mass"] == "1"     # group from childclass xyz_base: NOT in 4-5
    claw = {g.get("name"): g for g in by["rightclaw"]["geoms"]}["rightclaw_it"]
    assert claw["group"] == "4" and claw["contype"] == "1"                          # class base_col
    j = {x["name"]: x for b in spec.bodies for x in b["joints"]}
    assert j["right_j3"]["damping"] == "10" and j["right_j3"]["armature"] == "0.001" and j["right_j3"]["limited"] == "true"
    assert j["r_close"]["armature"] == "100" and j["r_close"]["damping"] == "1000" and j["doorjoint"]["damping"] == "2"
    assert by["mocap"]["mocap"] and len(spec.actuators) == 2 and spec.actuators[0]["kp"] == "400"
    raw = C.RawModel(spec)
    hand = raw.names.index("hand")
    assert raw.mass[hand] == 0.0 and np.allclose(raw.ipos[hand], [0, 0, 0.12])      # massless: ipos <- pos (MuJoCo compiler)
    assert abs(raw.mass[raw.names.index("rightclaw")] - 0.0162) < 1e-12             # 1000 kg/m^3 x the claw box
    assert raw.mass[raw.names.index("rightpad")] == 0.0                             # mass="1" geom is outside the group range
    assert raw.nq == 10 and raw.nv == 10
    assert abs(raw.mass[raw.names.index("door_link")] - 0.111491) < 1e-6         # five density-50 collision geoms (group 4)
    biw, diw = raw.invweight0()
    assert abs(biw[hand, 0] - 6.1056) < 1e-3 and abs(biw[hand, 1] - 284.895) < 1e-2


def test_legacy_mesh_centre_of_the_door_handle(CM):
    """MuJoCo 2.1's LEGACY mesh centring of door_handle.stl (the observed object position)."""
    from earl_benchmark_b200.mjcf import mesh
    tris = mesh.load_stl(os.path.join(REF, "earl_benchmark/envs/metaworld_assets/objects/meshes/doorlock/door_handle.stl"), np.ones(3))
    assert len(tris) == 896
    assert np.abs(mesh.legacy_center(tris) - np.array([5.07216292e-02, 2.594e-09, 4.51399102e-02])).max() < 5e-9
