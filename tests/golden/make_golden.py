#!/usr/bin/env python
"""Generate the committed golden fixtures (run once, in the build container).

The reference ships no tests, fixtures or golden vectors, so the
pins are made here from the two independent sources this container has:

  * ``taps_*.npz``      -- the Blackman windowed-sinc low-cut taps evaluated with
                           mpmath at 50 digits and rounded once to binary64
                           (independent of the oracle's x87 long double path);
  * ``filtercore_*.npz``-- outputs of the REFERENCE's own FilterCore.h
                           apply_filter_range(), compiled in place from
                           /root/reference against interface shims
                           (oracle/_ref/libref_filtercore.so, see
                           oracle/ref_filtercore.cpp): the index / edge / float32
                           narrowing logic that ran is the reference's.

Neither mpmath's presence nor /root/reference is needed to *run* the tests; only
this script needs them.  Usage:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))

import oracle  # noqa: E402


def mp_taps(fc, bw):
    import mpmath as mp

    mp.mp.dps = 50
    M = oracle.kernel_order(bw)
    H = M // 2
    fc = mp.mpf(fc)
    h = []
    for i in range(M + 1):
        m = i - H
        s = 2 * mp.pi * fc if m == 0 else mp.sin(2 * mp.pi * fc * m) / m
        w = mp.mpf("0.42") - mp.mpf("0.5") * mp.cos(2 * mp.pi * i / M) + mp.mpf("0.08") * mp.cos(4 * mp.pi * i / M)
        h.append(s * w)
    S = mp.fsum(h)
    out = [-(v / S) for v in h]
    out[H] += 1
    return np.array([float(v) for v in out], dtype=np.float64)


def main():
    # ---- taps ------------------------------------------------------------------
    for name, fs, f, s in [("small", 48000.0, 20.0, 200.0), ("cfg1", 48000.0, 20.0, 20.0),
                           ("odd", 44100.0, 440.0, 80.0), ("cfg2", 44100.0, 30.0, 10.0)]:
        fc, bw = f / fs, s / fs
        np.savez_compressed(os.path.join(HERE, f"taps_{name}.npz"), fc=fc, bw=bw, taps=mp_taps(fc, bw))
        print("taps", name, oracle.kernel_order(bw) + 1)

    # ---- FilterCore.h outputs ----------------------------------------------------
    assert oracle.ref_lib() is not None, "oracle/_ref was not built (needs /root/reference)"
    rng = np.random.default_rng(0xF1F1)
    cases = {}
    # (name, n_samples, fs, f, s): one long-enough signal, one shorter than 2*taps,
    # the "-f 440 -s 80" kernel of the reference's own `make test` (Makefile:47).
    for name, n, fs, f, s in [("body", 6000, 48000.0, 20.0, 200.0), ("maketest", 9000, 44100.0, 440.0, 80.0),
                              ("short", 1500, 48000.0, 20.0, 200.0)]:
        taps = oracle.build_lowcut(f / fs, s / fs)
        x = (rng.uniform(-0.9, 0.9, n) + 0.05).astype(np.float32)
        y_full = oracle.ref_fir_f32(x, taps)                     # one range [0, n)
        y_thr = oracle.ref_filter_channel_threads(x, taps, 5)    # ProcessFile.cp:60-83, 5 threads
        # a range that starts inside the prologue and ends inside the epilogue
        lo, hi = 7, n - 3
        y_rng = oracle.ref_fir_f32(x, taps, lo, hi)
        cases[name] = dict(x=x, taps=taps, y_full=y_full, y_thr=y_thr, y_rng=y_rng, lo=lo, hi=hi)
        assert np.array_equal(y_full, y_thr), "reference result depends on the thread partition?"
        print("filtercore", name, n, taps.size)
    for name, d in cases.items():
        np.savez_compressed(os.path.join(HERE, f"filtercore_{name}.npz"), **d)


if __name__ == "__main__":
    main()
