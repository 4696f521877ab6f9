"""Launch geometry of the skinning kernel for every output layout (dmk_skin_plan works without a device): the index
arithmetic of skin_kernel.cu -- v0 = tile * tile_vertices + thread * vertices_per_thread, active iff inside the tile and the
mesh -- must cover every vertex of the padded mesh exactly once, inside the launch bounds of each kernel shape."""
import numpy as np
import pytest

from darmok_b200 import build, capi

MAX_THREADS = {0: 256, 1: 320, 2: 448, 3: 256, 4: 256}


@pytest.fixture(scope="module", autouse=True)
def _built():
    build.build()


@pytest.mark.parametrize("layout", [0, 1, 2, 3, 4])
def test_every_vertex_is_owned_by_exactly_one_thread(layout):
    for num_vertices in [1, 7, 8, 9, 255, 1249, 1256, 1281, 1792, 1793, 2048, 2049, 2560, 2561, 5000, 6007, 20000, 65536]:
        vpad = (num_vertices + 7) // 8 * 8
        for num_sms in (148, 140, 3, 1):
            p = capi.skin_plan(vpad, 68, num_sms, layout)
            vpt = p["vertices_per_thread"]
            assert vpt == (4 if layout in (1, 2) else 8)
            assert p["threads"] % 32 == 0 and 32 <= p["threads"] <= MAX_THREADS[layout]
            assert p["tile_vertices"] % 8 == 0 and p["num_tiles"] * p["tile_vertices"] >= vpad
            assert p["threads"] * vpt >= p["tile_vertices"]
            # the persistent grid: never more blocks than SMs (one block per SM without a device).  Layouts 1 and 2 only exist in the
            # round-1 kernel shape, whose static split wants whole groups of tile slots; the producer-warp kernel takes characters from
            # per-tile counters and uses every SM
            tiles_per_pass = min(p["num_tiles"], p["grid"])
            assert 1 <= p["grid"] <= num_sms
            assert p["grid"] % tiles_per_pass == 0 if layout in (1, 2) else p["grid"] == num_sms
            owner = np.zeros(vpad, np.int32)
            for tile in range(p["num_tiles"]):
                tid = np.arange(p["threads"])
                v0 = tile * p["tile_vertices"] + tid * vpt
                active = (tid * vpt < p["tile_vertices"]) & (v0 < vpad)
                for k in range(vpt):
                    assert (v0[active] + k < vpad).all()
                    np.add.at(owner, v0[active] + k, 1)
            assert (owner == 1).all(), (layout, num_vertices, num_sms)


def test_large_palettes_fall_back_to_fewer_replicas_and_fit_shared_memory():
    for ps in (1, 68, 181, 226, 227, 256):
        p = capi.skin_plan(5000, ps, 148, 0)
        assert p["smem"] <= 227 * 1024
        assert p["replicas"] == (8 if ps <= 226 else 4)


def test_plan_rejects_bad_arguments():
    out = (capi.C.c_int64 * 7)()
    assert capi.lib().dmk_skin_plan(5001, 68, 148, 0, out) == capi.ERR_INVALID     # pitch not a multiple of 8
    assert capi.lib().dmk_skin_plan(5000, 68, 148, 5, out) == capi.ERR_INVALID     # unknown layout
    assert capi.lib().dmk_skin_plan(5000, 300, 148, 0, out) == capi.ERR_INVALID
