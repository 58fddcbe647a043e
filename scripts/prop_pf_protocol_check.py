"""CPU check of the stage protocol of the propagation kernel variants (csrc/gnnb_prop_body.cuh): the sequence of (item, chunk,
half) stages the PREFETCH gather loop fills must equal the sequence the MMA warp consumes and the default gather loop fills,
for any plan and any number of CTAs — a mismatch would hang the mbarrier pipeline.  Python transliteration of the three loops."""
import random


def mma_sequence(rank, nranks, nitems, ntiles, tile_chunk0, ksteps):
    seq = []
    item = rank
    while item < nitems:
        tile = item % ntiles
        for ch in range(tile_chunk0[tile], tile_chunk0[tile + 1]):
            nks = ksteps[ch]
            h = 0
            while 2 * h < nks:
                seq.append((item, ch, h)); h += 1
        item += nranks
    return seq


gather_default = mma_sequence          # the default gather loop has the same loop nest


def gather_prefetch(rank, nranks, nitems, ntiles, tile_chunk0, ksteps):
    seq = []
    item = rank
    if item >= nitems:
        return seq
    ch, ch1 = tile_chunk0[item % ntiles], tile_chunk0[item % ntiles + 1]
    nks = ksteps[ch]
    while True:
        nitem, nch, nch1 = item, ch + 1, ch1
        if nch >= ch1:
            nitem = item + nranks
            if nitem < nitems:
                nt = nitem % ntiles
                nch, nch1 = tile_chunk0[nt], tile_chunk0[nt + 1]
        has_next = nitem < nitems
        nnks = ksteps[nch] if has_next else 0
        h = 0
        while 2 * h < nks:
            seq.append((item, ch, h)); h += 1
        if not has_next:
            break
        item, ch, ch1, nks = nitem, nch, nch1, nnks
    return seq


random.seed(0)
for trial in range(2000):
    ntiles = random.randint(1, 9)
    chunks = [random.randint(1, 6) for _ in range(ntiles)]
    tile_chunk0 = [0]
    for c in chunks:
        tile_chunk0.append(tile_chunk0[-1] + c)
    ksteps = [random.randint(1, 4) for _ in range(tile_chunk0[-1])]
    ngroups = random.randint(1, 7)
    nitems = ntiles * ngroups
    nranks = random.randint(1, 12)
    for rank in range(nranks):
        a = mma_sequence(rank, nranks, nitems, ntiles, tile_chunk0, ksteps)
        b = gather_prefetch(rank, nranks, nitems, ntiles, tile_chunk0, ksteps)
        assert a == b, (trial, rank, a[:5], b[:5])
print('ok: the prefetch gather loop fills exactly the stages the MMA warp consumes (2000 random plans)')
