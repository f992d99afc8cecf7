"""CPU tests of the host-side planning of a level's cut search (orb_plan_level: pure host arithmetic, the same functions
orb_build uses).  What matters most: with several ranks every quantity that shapes an exchange (bins, slot size, which
search a level uses, whether the partition pre-builds the rows) depends only on numbers all ranks share - a rank that
took a different branch would leave its peers waiting in a cross-rank barrier or a collective."""
import pytest

POW2 = {1 << k for k in range(5, 21)}


def fields(p):
    return (p.search, p.hist_bins, p.cand_cap, p.slot_words, p.hist_words, p.prefuse_bins)


@pytest.mark.parametrize("ranks", [2, 4, 8])
@pytest.mark.parametrize("x_min", [14, 20, 24, 27])
def test_multi_rank_plan_depends_on_shared_numbers_only(orb, ranks, x_min):
    """Peer-memory protocol (orb_exchange.cuh): every level of every tree takes the selection search - streaming passes
    (3), one block per cell (4) or one warp per cell (5) - and all ranks plan it alike."""
    n_min = (1 << x_min) + 12345
    shards = [n_min, n_min + n_min // 3, 2 * n_min][:ranks] + [n_min + 7] * max(0, ranks - 3)
    n_global = sum(shards)
    for y in (4, 12, 16, 20):
        slot_total = max(1 << 20, 64 << y, ((n_min // 8) + 63) & ~63)
        for level in range(1, y + 1):
            n_cells = 1 << (level - 1)
            plans = [fields(orb.plan_level(n, n_cells, 1 << y, n_ranks=ranks, n_global=n_global, n_local_min=n_min)) for n in shards]
            assert all(p == plans[0] for p in plans), (ranks, x_min, y, level, plans)
            search, bins, cand_cap, slot_words, hist_words, pre = plans[0]
            hist_fit = n_min // 8 + 2 * 8192                     # fits every rank's buffer, also the smallest shard's
            if search == 0:      # only when even 32 bins per cell do not fit the smallest shard's rows
                assert n_cells * 32 > hist_fit or 32 * n_cells > slot_total
                assert (bins, cand_cap, slot_words, hist_words, pre) == (0, 0, 0, 0, 0)
                continue
            assert hist_words == n_cells * bins and hist_words <= hist_fit
            assert slot_words in POW2 and slot_words * n_cells <= slot_total
            assert cand_cap <= 49152
            lavg = max(n_global // ranks, n_min) // n_cells
            if search == 3:
                assert (n_cells < 512 or lavg > 1 << 20) and lavg > 8192
                assert bins in (512, 1024, 2048, 4096, 8192)
                assert pre in (0, 512, 1024) and (pre == 0 or pre == bins)
            else:
                assert search in (4, 5) and (n_cells >= 512 or lavg <= 8192) and pre == 0
                assert bins in (32, 64, 128, 256, 512, 1024, 2048)
                assert (search == 5) == (lavg < 4096)


def test_north_star_levels_never_take_the_iterative_search(orb):
    """C3 (2^27 -> 2^16) on 2, 4, 8 ranks and C5 (2^30 -> 2^20) on 8 ranks: every level uses the selection search."""
    for x, y, ranks in ((27, 16, 2), (27, 16, 4), (27, 16, 8), (30, 20, 8)):
        n = (1 << x) // ranks
        kinds = [orb.plan_level(n, 1 << (l - 1), 1 << y, n_ranks=ranks).search for l in range(1, y)]
        assert all(k in (3, 4, 5) for k in kinds), (x, y, ranks, kinds)
        assert kinds[0] == 3 and kinds[-1] in (4, 5)


@pytest.mark.parametrize("x", [10, 16, 20, 24, 26, 27, 30])
def test_single_rank_plan_bounds(orb, x):
    n = 1 << x
    for y in (3, 12, 16, 20):
        for level in range(1, y + 1):
            n_cells = 1 << (level - 1)
            for mode in (-1, 0, 1):
                p = orb.plan_level(n, n_cells, 1 << y, prefuse=mode)
                assert p.search in (1, 2) and p.slot_words == 0
                assert p.hist_words <= n // 16 + 2 * 8192
                assert p.cand_cap <= 49152
                assert p.prefuse_bins in (0, 512, 1024)
                if mode == 0 or level == 1:
                    assert p.prefuse_bins == 0
                if p.prefuse_bins:
                    assert p.hist_bins == p.prefuse_bins and p.hist_words == n_cells * p.prefuse_bins
                if p.search == 1:
                    assert p.hist_bins in (512, 1024, 2048, 4096, 8192) and p.hist_words == n_cells * p.hist_bins
                    # a bin of the streaming levels holds at most 16384 particles on average: one block can stage a few bins
                    assert (n // n_cells) // p.hist_bins <= 16384 or p.hist_bins == 8192
                if p.search == 2:
                    assert n_cells >= 64 and n // n_cells <= 1 << 20


def test_automatic_prefuse_policy(orb):
    """Several ranks: the partition pre-builds the next level's rows wherever the level streams.  One rank: never by
    default - below 2^25 particles per GPU a HIST pass is cheaper than binning in the partition, from 2^25 on the rows
    come from a sample instead (test_sampled_rows_policy); ORB_PREFUSE=1 still forces it."""
    small = [orb.plan_level(1 << 24, 1 << l, 1 << 12).prefuse_bins for l in range(1, 11)]
    big = [orb.plan_level(1 << 27, 1 << l, 1 << 16).prefuse_bins for l in range(1, 15)]
    forced = [orb.plan_level(1 << 24, 1 << l, 1 << 12, prefuse=1).prefuse_bins for l in range(1, 11)]
    forced_big = [orb.plan_level(1 << 27, 1 << l, 1 << 16, prefuse=1).prefuse_bins for l in range(1, 15)]
    multi = [orb.plan_level(1 << 24, 1 << l, 1 << 12, n_ranks=2).prefuse_bins for l in range(1, 11)]
    assert not any(small) and not any(big)
    assert all(forced)
    assert forced_big[:2] == [0, 0] and all(forced_big[3:])   # 2^26 / 2^25-particle cells want more than 1024 bins: separate pass
    # 2 x 2^24 in 2 cells: 2048 bins wanted, more than the partition holds; from 512 cells on the cells are searched by
    # one block each (their rows come from k_xd_hist, not from the partition)
    assert multi[0] == 0 and all(multi[1:8]) and not any(multi[8:])


def test_sampled_rows_policy(orb):
    """Sampled histogram rows: one rank, builds of at least 2^25 particles per GPU; streaming levels
    (with the one-block FINISH, ORB_PAR_FINISH=0, only where cells hold at most 2^25 particles), block-searched cells of
    at least 2^16; never with several ranks, never
    where the partition-built rows are forced."""
    c3 = [orb.plan_level(1 << 27, 1 << l, 1 << 16) for l in range(0, 15)]
    assert [p.sample_stride for p in c3] == [8] * 12 + [1, 1, 1]
    assert [p.search for p in c3] == [1] * 7 + [2] * 8
    # the sampled streaming levels bin finely (the margin, not the bin width, sets the number of candidates)
    assert all(p.hist_bins == 8192 for p in c3[:5]) and all(p.hist_words == (1 << l) * p.hist_bins for l, p in enumerate(c3[:7]))
    assert all(p.cand_cap <= 49152 for p in c3)
    c2 = [orb.plan_level(1 << 24, 1 << l, 1 << 12) for l in range(0, 11)]
    assert all(p.sample_stride == 1 for p in c2)
    c4 = [orb.plan_level(1 << 26, 1 << l, 1 << 14) for l in range(0, 13)]
    assert [p.sample_stride for p in c4] == [8] * 11 + [1, 1]
    assert all(orb.plan_level(1 << 27, 1 << l, 1 << 16, n_ranks=2).sample_stride <= 1 for l in range(0, 15))
    assert all(orb.plan_level(1 << 27, 1 << l, 1 << 16, prefuse=1).sample_stride <= 1 for l in range(0, 15))
