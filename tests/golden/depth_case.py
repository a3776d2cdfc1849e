"""A targeted-mode case on which the reference's depth cut (libkit4b/SfxArray.cpp:4480-4494) fires.

K=32, R=3: the last cascade level looks up the four 8-base cores of a probe K-mer and walks at most
4 x MaxCoreDepth suffix-array entries per core (hammings.cpp:2366-2386; multiplier SfxArray.cpp:4303).
The assembly holds 42000 copies of one 8-mer in random flanks, all sorting BEFORE the single locus that
is 3 mismatches away from the probe K-mer U - one mismatch in each half and each third of U, so that
only core 2 (that 8-mer) is intact and the earlier cascade levels cannot find the locus.  With
-s0 (default, 40000 entries) and -s3 (20000) the reference gives the core up and reports 4 ("not
found"); with -s1 / -s2 it walks far enough and reports the true minimum 3.  make_depth_golden.py ran
the unmodified reference on exactly these sequences; the tests rebuild them from the seed."""
import numpy as np

K, R = 32, 3


def build():
    """-> (assembly entries [(name, codes)], probe entries [(name, codes)], position of U in the first probe)"""
    rng = np.random.default_rng(2024)
    u = rng.integers(0, 4, size=K, dtype=np.uint8)
    u[24:28] = 3                      # what follows core 2 at the true locus sorts last among the copies of core 2
    n_copies = 42000
    fl = rng.integers(0, 4, size=(n_copies, 16), dtype=np.uint8)
    fl[:, 8] = rng.integers(0, 3, size=n_copies)   # copies are followed by A/C/G: they sort before the locus
    block = np.concatenate([fl[:, :8], np.tile(u[16:24], (n_copies, 1)), fl[:, 8:]], axis=1).ravel()
    bg = rng.integers(0, 4, size=200_000, dtype=np.uint8)
    locus = u.copy()
    for p in (5, 15, 25):             # one mismatch per third and per half; only core 2 stays intact
        locus[p] = (locus[p] + 1) % 4 if p != 25 else 2
    chr_a = np.concatenate([bg[:100_000], block, bg[100_000:]]).astype(np.uint8)
    chr_b = np.concatenate([rng.integers(0, 4, size=3000, dtype=np.uint8), locus,
                            rng.integers(0, 4, size=3000, dtype=np.uint8)]).astype(np.uint8)
    probes = [("pU", np.concatenate([rng.integers(0, 4, size=10, dtype=np.uint8), u,
                                     rng.integers(0, 4, size=10, dtype=np.uint8)]).astype(np.uint8)),
              ("rnd", rng.integers(0, 4, size=200, dtype=np.uint8))]
    return [("chrA", chr_a), ("chrB", chr_b)], probes, 10
