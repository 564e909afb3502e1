"""Host-side restatement of the tile geometry of the row-blocked kernels (csrc/substage_rb.cu): every
cell of a launch is owned by exactly one (tile, warp, lane, row), every stencil stays inside the TMA box,
and the box of the divergence kernel starts at an even FP64 column (an odd start column faults on B200).
The constants are parsed from the source so that the test follows the kernel."""
import re
from pathlib import Path

import pytest

SRC = (Path(__file__).resolve().parent.parent / "swmhd_b200" / "csrc" / "substage_rb.cu").read_text()


def _macro(name):
    m = re.search(rf"#define {name} (\d+)", SRC)
    assert m, name
    return int(m.group(1))


TX = 32
TYB = _macro("RB_TY")
R = _macro("RB_R")
NW = TYB // R
W, HT = TX + 6, TYB + 6            # TMA box (doubles)
TXD = TX - 1                       # cells per tile row of the divergence kernel


def test_constants_match_the_source():
    assert "constexpr int TX = 32, TYB = RB_TY, R = RB_R, NW = TYB / R, NT = 32 * NW;" in SRC
    assert "constexpr int TXD = TX - 1;" in SRC
    assert "constexpr int W = TX + 6, HT = TYB + 6" in SRC
    assert TYB % 8 == 0 and TYB % R == 0


@pytest.mark.parametrize("form", ["jacobian", "divergence"])
@pytest.mark.parametrize("Nx", [8, 31, 32, 33, 62, 63, 64, 100, 1024])
@pytest.mark.parametrize("rows", [(0, 8), (0, 24), (8, 40), (16, 19), (0, 100)])
def test_every_cell_is_owned_exactly_once(form, Nx, rows):
    row_begin, row_end = rows
    cells_x = TX if form == "jacobian" else TXD
    tiles_x = (Nx + cells_x - 1) // cells_x
    tiles_y = (row_end - row_begin + TYB - 1) // TYB
    seen = {}
    for tile in range(tiles_x * tiles_y):
        tile_x, tile_y = tile % tiles_x, tile // tiles_x
        row0 = row_begin + tile_y * TYB
        xoff = tile_x * cells_x
        c0 = xoff & ~1 if form == "divergence" else xoff          # first parent column of the box
        assert c0 % 2 == 0, "TMA box must start at an even FP64 column"
        shift = xoff - c0
        for wp in range(NW):
            for lane in range(32):
                li = lane + 3 + shift
                i = xoff + 1 + lane                                   # logical 1-based column
                own = i <= Nx and (form == "jacobian" or lane < TXD)
                assert c0 + li == i + 2, "tile-local column must address the cell's parent column"
                # face-based WENO5 stencils: left-biased f-3..f+1, right-biased f-2..f+2, f = li
                assert li - 3 >= 0 and li + 2 <= W - 1, "x stencils (li-3 .. li+2) stay inside the box"
                for r in range(R):
                    lj = 3 + wp * R + r
                    assert lj - 3 >= 0 and lj + 3 <= HT - 1, "y stencils of the south and north faces stay inside the box"
                    j = row0 + 1 + wp * R + r                         # logical 1-based row
                    if own and j <= row_end:
                        assert (i, j) not in seen, f"cell {(i, j)} owned twice"
                        seen[(i, j)] = tile
    want = {(i, j) for i in range(1, Nx + 1) for j in range(row_begin + 1, row_end + 1)}
    assert set(seen) == want


def test_diag_partial_slots_are_unique_per_launch_granule():
    """A 16-row tile starting at 8-row granule tr8 fills slot (tr8, tile_x) and neutral elements into the
    other granules it covers: slots never collide between the edge and interior launches of a slab."""
    Nx, Ny = 100, 72
    ntr = (Ny + 7) // 8
    for form, cells_x in (("jacobian", TX), ("divergence", TXD)):
        tiles_x = (Nx + cells_x - 1) // cells_x
        written = {}
        for row_begin, row_end in ((0, 8), (8, Ny - 8), (Ny - 8, Ny)):      # edges + interior of a slab substage
            tiles_y = (row_end - row_begin + TYB - 1) // TYB
            for tile_y in range(tiles_y):
                row0 = row_begin + tile_y * TYB
                for tile_x in range(tiles_x):
                    for s in range(TYB // 8):
                        if s == 0 or row0 + 8 * s < row_end:
                            slot = (row0 // 8 + s) * tiles_x + tile_x
                            assert slot not in written, (form, slot)
                            written[slot] = (row_begin, tile_y, tile_x, s)
        assert sorted(written) == list(range(tiles_x * ntr))
