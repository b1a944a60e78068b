#!/usr/bin/env python3
"""GPU debug aid: dump one 128x128 tcgen05 distance tile with several UMMA
descriptor stride settings and compare with exact integer arithmetic."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from imageanalysis_b200 import _capi, synth  # noqa: E402


def main():
    out = {}
    for norm, gen, nb, name in ((_capi.NORM_L2, synth.sift_like, 128, "l2"), (_capi.NORM_HAMMING, synth.orb_like, 32, "ham")):
        q, t = gen(300, seed=1), gen(300, seed=2)
        eng = _capi.Engine(norm, nb, 0)
        eng.upload(0, q)
        eng.upload(1, t)
        if norm == _capi.NORM_L2:
            exp = ((q[:128, None].astype(np.int64) - t[None, :128].astype(np.int64)) ** 2).sum(-1)
            dot = -2 * (q[:128].astype(np.int64) @ t[:128].astype(np.int64).T)
        else:
            # the accumulator is the key 32 * distance + (train row mod 32); without the augmentation K-step: -64 q.t
            exp = 32 * np.unpackbits(q[:128, None] ^ t[None, :128], axis=-1).sum(-1).astype(np.int64) + (np.arange(128) % 32)[None, :]
            dot = -64 * (np.unpackbits(q[:128], axis=-1).astype(np.int64) @ np.unpackbits(t[:128], axis=-1).astype(np.int64).T)
        for tag, kw in (("prod", dict()), ("prod_A_in_tmem", dict(ksteps=-9)), ("data_only", dict(ksteps=8)),
                        ("data_only_A_in_tmem", dict(ksteps=-8))):
            try:
                got = eng.debug_tile(0, 1, **kw)
            except Exception as e:  # noqa: BLE001
                print(name, tag, "FAILED:", e)
                continue
            ref = exp if tag.startswith("prod") else dot
            diff = np.abs(got.astype(np.float64) - ref)
            print("%s %-11s max|diff|=%g  exact=%d/16384  got[0,:4]=%s ref[0,:4]=%s" % (
                name, tag, diff.max(), int((diff == 0).sum()), got[0, :4].tolist(), ref[0, :4].tolist()))
            out["%s_%s" % (name, tag)] = got
        out[name + "_exp"] = exp
        if norm == _capi.NORM_L2:
            # byte layout (kind::i8): rows in rank order (even squared norms first), acc = q.t + CAP - floor(|t|^2/2)
            CAP = 254 + 255 * 254 + 30 * 65025

            def ranked(x):
                nrm = (x.astype(np.int64) ** 2).sum(1)
                order = np.concatenate([np.flatnonzero(nrm % 2 == 0), np.flatnonzero(nrm % 2 == 1)])
                return x[order].astype(np.int64), nrm[order]
            qr, _ = ranked(q)
            tr, tn = ranked(t)
            ref8 = qr[:128] @ tr[:128].T + CAP - tn[None, :128] // 2
            for tag, kw in (("i8_A_in_tmem", dict(lbo=0, ksteps=-5)), ("i8_A_in_smem", dict(lbo=0, ksteps=5))):
                try:
                    got = eng.debug_tile(0, 1, **kw)
                except Exception as e:  # noqa: BLE001
                    print(name, tag, "FAILED:", e)
                    continue
                diff = np.abs(got.astype(np.float64) - ref8)
                print("%s %-11s max|diff|=%g  exact=%d/16384  got[0,:4]=%s ref[0,:4]=%s" % (
                    name, tag, diff.max(), int((diff == 0).sum()), got[0, :4].tolist(), ref8[0, :4].tolist()))
                out["%s_%s" % (name, tag)] = got
            out["l2_i8_exp"] = ref8
        eng.close()
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    np.savez_compressed(os.path.join(ROOT, "gpurun_out", "debug_tile.npz"), **out)


if __name__ == "__main__":
    main()
