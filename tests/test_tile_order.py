"""Host-side restatement of conv_umma.cuh::tile_coords for the three tile orders (N tile fastest, pixel tile fastest, CTA
pairs): every (N tile, pixel tile) is visited exactly once, and in pair mode the two CTAs of a cluster -- tiles t and t+1
with t even, walked in lockstep by CTAs 2P and 2P+1 of a persistent grid with an even number of CTAs -- always hold the
same N tile on adjacent pixel tiles (what tcgen05.mma.cta_group::2 needs: one weight tile, M = 256 rows)."""
import itertools

import pytest


def tile_coords(t, n_tiles, m_total, pair=False, m_fast=False):
    if pair:
        q, r = t >> 1, t & 1
        mp = q // n_tiles
        return q - mp * n_tiles, 2 * mp + r
    if m_fast:
        nt = t // m_total
        return nt, t - nt * m_total
    mt = t // n_tiles
    return t - mt * n_tiles, mt


@pytest.mark.parametrize("n_tiles,m_total", [(1, 2), (3, 256), (8, 32), (2, 4096), (4, 6)])
def test_every_tile_once(n_tiles, m_total):
    total = n_tiles * m_total
    want = set(itertools.product(range(n_tiles), range(m_total)))
    for kw in (dict(), dict(m_fast=True), dict(pair=True)):
        got = [tile_coords(t, n_tiles, m_total, **kw) for t in range(total)]
        assert len(set(got)) == total and set(got) == want, kw


@pytest.mark.parametrize("n_tiles,m_total,grid", [(3, 256, 148), (1, 4096, 148), (8, 32, 148), (4, 6, 24), (2, 2, 4)])
def test_pairs_walk_in_lockstep(n_tiles, m_total, grid):
    total = n_tiles * m_total
    assert total % 2 == 0 and grid % 2 == 0
    for cta in range(0, min(grid, total), 2):
        mine = list(range(cta, total, grid))
        peer = list(range(cta + 1, total, grid))
        assert len(mine) == len(peer)  # both CTAs of a cluster run the same number of tiles (no one-sided waits)
        for t0, t1 in zip(mine, peer):
            (n0, m0), (n1, m1) = tile_coords(t0, n_tiles, m_total, pair=True), tile_coords(t1, n_tiles, m_total, pair=True)
            assert n0 == n1 and m1 == m0 + 1 and m0 % 2 == 0
