"""Property tests of the launch-partition arithmetic of the Marlin kernels, restated in Python line by line.

Small-batch kernel (csrc/marlin_gemm_small.cu: `small_partition` on the host, the `u / cta_of / nseg / slab index`
arithmetic in the kernel): for every shape and SM count each (tile, chunk) unit is processed exactly once, every CTA
contributes at most one segment per tile, slab indices are unique per tile and below the bound `b200_marlin_gemm_plan`
reports (the size of the caller's fp32 scratch), and the ticket count a tile waits for equals the number of CTAs that touch
it. tcgen05 kernel (csrc/marlin_gemm.cu): the k-split actually launched never exceeds the planned bound, never leaves
a split empty and keeps all splits co-resident. Grouped (MoE) launch: the upper-bound grid enumerates every sorted row
exactly once."""
import random

import pytest


def small_partition(N, K, sms):                      # host: grid size and slab bound
    tiles, chunks = (N + 127) // 128, K // 64
    U = tiles * chunks
    G = min(2 * sms, max(1, U // 8))
    slabs = 1
    for t in range(tiles):
        cf = ((t * chunks + 1) * G - 1) // U
        cl = ((t * chunks + chunks) * G - 1) // U
        slabs = max(slabs, cl - cf + 1)
    return G, slabs


def kernel_segments(N, K, G):                        # device: what every CTA does
    tiles, chunks = (N + 127) // 128, K // 64
    U = tiles * chunks
    cta_of = lambda u: ((u + 1) * G - 1) // U
    out = []
    for c in range(G):
        u, u_end = c * U // G, (c + 1) * U // G
        while u < u_end:
            tile = u // chunks
            cb = u - tile * chunks
            ce = min(chunks, cb + (u_end - u))
            u += ce - cb
            cf = cta_of(tile * chunks)
            nseg = cta_of(tile * chunks + chunks - 1) - cf + 1
            out.append((c, tile, cb, ce, nseg, c - cf))
    return out


def _check(N, K, sms):
    tiles, chunks = (N + 127) // 128, K // 64
    G, slabs = small_partition(N, K, sms)
    assert 1 <= G <= 2 * sms
    segs = kernel_segments(N, K, G)
    covered = [[0] * chunks for _ in range(tiles)]
    per_tile = {}
    for c, tile, cb, ce, nseg, slab in segs:
        assert 0 <= cb < ce <= chunks
        for k in range(cb, ce):
            covered[tile][k] += 1
        assert 0 <= slab < nseg <= slabs, (N, K, sms, c, tile, slab, nseg, slabs)
        per_tile.setdefault(tile, []).append((c, slab, nseg))
    assert all(v == 1 for row in covered for v in row), (N, K, sms)
    for tile, lst in per_tile.items():
        nseg = lst[0][2]
        assert all(n == nseg for _, _, n in lst)
        assert len(lst) == nseg                                   # tickets awaited == CTAs that arrive
        assert sorted(s for _, s, _ in lst) == list(range(nseg))  # slab indices unique and dense
        assert len({c for c, _, _ in lst}) == nseg                # one segment per CTA and tile
    work = {}
    for c, tile, cb, ce, _, _ in segs:
        work[c] = work.get(c, 0) + ce - cb
    assert max(work.values()) - min(work.values()) <= 1            # equal shares (stream-k)


@pytest.mark.parametrize("N,K", [(6144, 4096), (4096, 4096), (28672, 4096), (4096, 14336), (64, 64), (64, 128),
                                 (128, 64), (192, 8192), (128256, 4096), (1024, 28672), (8192, 8192), (320, 1024)])
@pytest.mark.parametrize("sms", [148, 132, 64])
def test_llama_and_edge_shapes(N, K, sms):
    _check(N, K, sms)


def test_random_shapes():
    rng = random.Random(7)
    for _ in range(300):
        _check(64 * rng.randint(1, 700), 64 * rng.randint(1, 300), rng.choice([148, 144, 132, 108, 80, 1]))


# ---- tcgen05 kernel: k-split plan (csrc/marlin_gemm.cu plan_split_k + the rounding in b200_gptq_marlin_gemm) ----------
def _plan_split_k(M, N, K, group_size, sms):
    tiles = ((N + 127) // 128) * ((M + 255) // 256)
    chunks = K // 64
    if (M + 255) // 256 > 32:
        return 1
    split = max(1, min(sms // max(tiles, 1), chunks // 8, 8))
    gchunks = group_size // 64 if group_size > 64 else 1
    while split > 1:
        per = -(-(-(-chunks // split)) // gchunks) * gchunks
        if (split - 1) * per < chunks:
            break
        split -= 1
    return split


def _launch_split(M, N, K, group_size, sms):
    split = _plan_split_k(M, N, K, group_size, sms)
    tiles = ((N + 127) // 128) * ((M + 255) // 256)
    chunks = K // 64
    while split > 1 and tiles * split > sms:
        split -= 1
    while split > 1 and (split - 1) * (-(-chunks // split)) >= chunks:
        split -= 1
    gchunks = group_size // 64 if group_size > 64 else 1
    per = -(-(-(-chunks // split)) // gchunks) * gchunks
    while split > 1 and (split - 1) * per >= chunks:
        split -= 1
        per = -(-(-(-chunks // split)) // gchunks) * gchunks
    return split, per


def test_k_split_plan_of_the_tcgen05_kernel():
    rng = random.Random(3)
    shapes = [(256, 6144, 4096, 128), (256, 4096, 4096, 128), (256, 28672, 4096, 128), (256, 4096, 14336, 128),
              (64, 4096, 4096, -1), (200, 64, 64, 32), (4096, 4096, 4096, 128), (9000, 128, 8192, 128)]
    shapes += [(rng.randint(33, 2000), 64 * rng.randint(1, 500), 64 * rng.randint(1, 256), rng.choice([-1, 32, 64, 128, 256]))
               for _ in range(300)]
    for M, N, K, gs in shapes:
        if gs > 0 and K % gs:
            continue
        for sms in (148, 132):
            plan = _plan_split_k(M, N, K, gs, sms)
            split, per = _launch_split(M, N, K, gs, sms)
            chunks = K // 64
            tiles = ((N + 127) // 128) * ((M + 255) // 256)
            assert 1 <= split <= plan                               # the caller's scratch [plan, M, N] is large enough
            assert split == 1 or (tiles * split <= sms and split <= 8)   # one wave of clusters of <= 8 CTAs
            assert (split - 1) * per < chunks <= split * per        # all chunks covered, no empty split
            if gs > 64:
                assert per % (gs // 64) == 0 or split == 1


# ---- grouped (MoE) launch: blockIdx.y -> (expert, tile) over an upper-bound grid ----------------------------------------
def test_moe_tile_enumeration_covers_every_sorted_row():
    rng = random.Random(5)
    for _ in range(300):
        E = rng.randint(1, 64)
        topk = rng.randint(1, min(E, 8))
        M = rng.randint(1, 700)
        block = rng.choice([16, 32, 64])
        counts = [0] * E
        for _t in range(M):
            for e in rng.sample(range(E), topk):
                counts[e] += 1
        offsets = [0]
        for c in counts:
            offsets.append(offsets[-1] + -(-c // block) * block)
        cap = M * topk + E * (block - 1)                             # sorted_ids allocation of the caller
        assert offsets[-1] <= cap
        blk = max(block, 16)
        tile_rows = min(256, (-(-M // blk) * blk + 15) // 16 * 16)
        max_tiles = E + cap // tile_rows
        seen = [0] * offsets[-1]
        for y in range(max_tiles):                                   # the kernel's enumeration
            t, e = y, 0
            while e < E:
                ln = offsets[e + 1] - offsets[e]
                nt = -(-ln // tile_rows)
                if t < nt:
                    break
                t -= nt
                e += 1
            if e >= E:
                continue
            base = offsets[e] + t * tile_rows
            for r in range(base, base + min(tile_rows, offsets[e + 1] - offsets[e] - t * tile_rows)):
                seen[r] += 1
        assert all(v == 1 for v in seen)


def test_python_restatement_equals_the_library_plan():
    """b200_marlin_gemm_plan runs without a GPU (148 SMs assumed): the restatements above are the shipped arithmetic."""
    from aphrodite_engine_b200 import _native
    lib = _native.load_c_abi()
    rng = random.Random(11)
    cases = [(256, 6144, 4096, 128), (256, 28672, 4096, 128), (256, 4096, 14336, 128), (16, 28672, 4096, 128),
             (1, 4096, 4096, -1), (32, 6144, 4096, 128), (33, 6144, 4096, 128), (7, 128256, 4096, 128)]
    cases += [(rng.randint(1, 600), 64 * rng.randint(1, 500), 64 * rng.randint(1, 256), rng.choice([-1, 32, 64, 128]))
              for _ in range(200)]
    for M, N, K, gs in cases:
        if gs > 0 and K % gs:
            continue
        groups = K // gs if gs > 0 else 1
        expect = _plan_split_k(M, N, K, gs if groups > 1 else -1, 148)
        if M <= 32:
            expect = max(expect, small_partition(N, K, 148)[1])
        assert lib.b200_marlin_gemm_plan(M, N, K, groups) == expect, (M, N, K, gs)


def test_scaled_mm_split_plan_properties():
    """W8A8 GEMM k-split plan (csrc/scaled_mm.cu, queried through the C ABI on a CPU box: no device, so the cluster-slot
    bound falls back to SMs / split): 1 <= split <= 8, never an empty split, at least 4 chunks (512 k) per split, never
    more CTAs than one wave, and the tile-shape override only changes the tile count it plans for."""
    from aphrodite_engine_b200 import _native
    lib = _native.load_c_abi()
    try:
        for tile in (0, 1, 2):
            lib.b200_scaled_mm_set_tile(tile)
            for M in (1, 16, 256, 300, 4096):
                for N in (64, 128, 4096, 6144, 28672):
                    for K in (128, 512, 1040, 4096, 14336):
                        split = lib.b200_scaled_mm_plan(M, N, K)
                        chunks = (K + 127) // 128
                        assert 1 <= split <= 8 and split <= max(1, chunks // 4)
                        per = (chunks + split - 1) // split
                        assert (split - 1) * per < chunks, (M, N, K, split)
        assert lib.b200_scaled_mm_plan(0, 128, 128) == 1 and lib.b200_scaled_mm_plan(16, 128, 0) == 1
    finally:
        assert lib.b200_scaled_mm_set_tile(0) in (0, 1, 2)
