"""Host model of the number representation the DP kernels use (telr_b200/csrc/k_fill.cuh, "offset form"): the packed 32-bit
arithmetic on offset half-words, the top-bit flags and the multiply-add flag gather are replayed with Python integers and
compared with the plain signed recurrence of ksw_extd2 for random cell inputs, for the scoring of every preset.  The kernels
themselves are checked bit for bit against the oracle in the GPU suite; this test pins the identities they rely on
(no borrow across bit 16, top bit == "is the maximum" / "gap continues", sum_f r_f 2^f == 255 T)."""
import random

import pytest

FB = 60
M32 = 0xFFFFFFFF
PRESETS = {"map-ont/map-pb": dict(a=2, b=4, q=4, e=2, q2=24, e2=1), "map-hifi": dict(a=1, b=4, q=6, e=2, q2=26, e2=1)}


def pk(lo, hi):
    assert 0 <= lo < 65536 and 0 <= hi < 65536
    return lo | hi << 16


def halves(w):
    return w & 0xFFFF, w >> 16


def s16(v):
    return v - 65536 if v & 0x8000 else v


def vimax3_s16x2(a, b, c):
    return pk(*[max(s16(x), s16(y), s16(z)) & 0xFFFF for x, y, z in zip(halves(a), halves(b), halves(c))])


def viaddmax_u16x2(a, b, c):
    return pk(*[max((x + y) & 0xFFFF, z) for x, y, z in zip(halves(a), halves(b), halves(c))])


def prmt_top(a, b):
    """prmt.b32 with selector 0xFDB9: bytes 1, 3 of a and of b, each replaced by 8 copies of its top bit."""
    bits = [(a >> 15) & 1, (a >> 31) & 1, (b >> 15) & 1, (b >> 31) & 1]
    return sum((0xFF if t else 0) << (8 * i) for i, t in enumerate(bits))


def plain_cell(sc, S, Lv, Lx, Lx2, up_u, up_y, up_y2):
    """One cell of the two-piece affine difference recurrence (left-aligned gaps), signed integers."""
    q, e, q2, e2 = sc["q"], sc["e"], sc["q2"], sc["e2"]
    A, A2, B, B2 = Lx + Lv, Lx2 + Lv, up_y + up_u, up_y2 + up_u
    Z = max(S, A, B, A2, B2)
    flags = (S < Z) | (A < Z) << 1 | (B < Z) << 2 | (A2 < Z) << 3
    flags |= (A - Z + q <= 0) << 4 | (B - Z + q <= 0) << 5 | (A2 - Z + q2 <= 0) << 6 | (B2 - Z + q2 <= 0) << 7
    return dict(nu=Z - Lv, nv=Z - up_u, nx=max(A - Z - e, -q - e), ny=max(B - Z - e, -q - e),
                nx2=max(A2 - Z - e2, -q2 - e2), ny2=max(B2 - Z - e2, -q2 - e2), flags=flags)


def offset_pair(sc, cells):
    """The same for two cells at once in the kernel's packed offset form; returns the two cells' new state and the packed words."""
    q, e, q2, e2 = sc["q"], sc["e"], sc["q2"], sc["e2"]
    qe, qe2 = q + e, q2 + e2
    BX, BX2 = qe - 1 + 0x8000, qe2 - 1 + 0x8000
    CA, CA2 = (FB - BX) * 65537 & M32, (FB - BX2) * 65537 & M32
    DK, XFLOOR = 0x80008000, 0x7FFF7FFF
    P = lambda key, off: pk(*[c[key] + off for c in cells])
    S, Lv, up_u = P("S", 2 * FB), P("Lv", FB), P("up_u", FB)
    Lx, up_y, Lx2, up_y2 = P("Lx", BX), P("up_y", BX), P("Lx2", BX2), P("up_y2", BX2)
    A, A2 = (Lx + Lv + CA) & M32, (Lx2 + Lv + CA2) & M32
    B, B2 = (up_y + up_u + CA) & M32, (up_y2 + up_u + CA2) & M32
    Z = vimax3_s16x2(vimax3_s16x2(S, A, B), A2, B2)
    NZ = (DK - Z) & M32
    DS, DA, DB, DA2, DB2 = [(t + NZ) & M32 for t in (S, A, B, A2, B2)]
    nu, nv = (Z - Lv) & M32, (Z - up_u) & M32
    nx, ny = viaddmax_u16x2(DA, pk(q - 1, q - 1), XFLOOR), viaddmax_u16x2(DB, pk(q - 1, q - 1), XFLOOR)
    nx2, ny2 = viaddmax_u16x2(DA2, pk(q2 - 1, q2 - 1), XFLOOR), viaddmax_u16x2(DB2, pk(q2 - 1, q2 - 1), XFLOOR)
    out = []
    for h in range(2):
        g = lambda w: halves(w)[h]
        out.append(dict(nu=g(nu) - FB, nv=g(nv) - FB, nx=g(nx) - BX, ny=g(ny) - BX, nx2=g(nx2) - BX2, ny2=g(ny2) - BX2))
    return out, (DS, DA, DB, DA2, nx, ny, nx2, ny2)


def rand_cell(sc, rng):
    qe, qe2, a, b = sc["q"] + sc["e"], sc["q2"] + sc["e2"], sc["a"], sc["b"]
    return dict(S=rng.choice((a, -b)), Lv=rng.randint(-qe2, a + qe2), up_u=rng.randint(-qe2, a + qe2),
                Lx=rng.randint(-qe, a + qe2), up_y=rng.randint(-qe, a + qe2), Lx2=rng.randint(-qe2, a + qe2), up_y2=rng.randint(-qe2, a + qe2))


@pytest.mark.parametrize("name", sorted(PRESETS))
def test_offset_form_equals_the_plain_recurrence(name):
    sc, rng = PRESETS[name], random.Random(7)
    assert sc["a"] + sc["b"] + max(sc["q"] + sc["e"], sc["q2"] + sc["e2"]) <= FB and 2 * FB + sc["a"] <= 127     # the fast path's admission test
    for _ in range(4000):
        quad = [rand_cell(sc, rng) for _ in range(4)]                 # (k lo, k hi), (k+1 lo, k+1 hi): one flag word
        want = [plain_cell(sc, **c) for c in quad]
        got0, w0 = offset_pair(sc, quad[:2])
        got1, w1 = offset_pair(sc, quad[2:])
        for g, w in zip(got0 + got1, want):
            assert g == {k: w[k] for k in g}
        # flag gather: one top-bit PRMT per flag, a tree of multiply-adds, one multiply
        r = [prmt_top(x, y) for x, y in zip(w0, w1)]
        two, four, sixteen = 2, 4, 16
        t = [(r[2 * i + 1] * two + r[2 * i]) & M32 for i in range(4)]
        acc = (((t[3] * four + t[2]) & M32) * sixteen + ((t[1] * four + t[0]) & M32)) & M32
        word = (acc * 0x01010101 - 1) & M32
        assert [(word >> (8 * i)) & 0xFF for i in range(4)] == [w["flags"] for w in want]


def test_a_garbage_high_half_cannot_disturb_the_low_half():
    """The first and last column iteration of a lane-step carry one meaningless cell; it always sits in the half a carry or
    borrow of the live half's arithmetic cannot come from (carries only travel upwards)."""
    sc, rng = PRESETS["map-ont/map-pb"], random.Random(11)
    for _ in range(2000):
        live = rand_cell(sc, rng)
        junk = {k: rng.randint(-FB, 200) for k in live}                # anything representable: raw zeros, stale values
        junk["S"] = rng.choice((sc["a"], -sc["b"]))
        got, _ = offset_pair(sc, [live, junk])
        want = plain_cell(sc, **live)
        assert got[0] == {k: want[k] for k in got[0]}
