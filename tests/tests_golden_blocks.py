"""Block boundaries of the flat [uw; vw; wT] gradient (shared by tests/golden/make_golden_bptt_long.py and the GPU test)."""


def blocks(P, sizes):
    out = []
    for net in range(3):
        off = net * P
        for i in range(len(sizes) - 1):
            for n_el in (sizes[i] * sizes[i + 1], sizes[i + 1]):
                out.append((off, off + n_el))
                off += n_el
    return out
