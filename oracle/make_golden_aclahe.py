#!/usr/bin/env python3
"""Golden vectors of the automatic CLAHE parameter search - TEST INFRASTRUCTURE, run in the build
container only (it imports the reference from /root/reference, which does not exist on the GPU box).

Imports the reference's own modules/aclahe/python/{ACLAHE,functions}.py (matplotlib is absent: stubbed, the plots are
side effects) and cv2 4.13.0, runs them on the reference's crowd.png and stores under tests/golden/crowd_full.npz:

  img            the 600x800 grey pixels `cv2.imread('crowd.png', 0)` hands ParametrosACLAHE (aclahe/python/main.py:17)
  as_committed   (BS, CL) returned by ACLAHE.ParametrosACLAHE exactly as the file is in the reference: the sweep body lost
                 its indentation (ACLAHE.py:40-47), so one CLAHE per block size is evaluated and the answer is (8, 0)
  repaired       (BS, CL) with the loop body indented the way the comments describe (every clip limit of every block
                 size evaluated), everything else - GaussianBlur, CLAHE, Entropia, graficar, DerivadaY / DerivadaX /
                 Curvatura, the float16 table of the block-size pick - being the reference's own code: (4, 7)
  entropies      [5][50] float32: Entropia(CLAHE(blur, BS, CL)) for BS in 2,4,8,16,32 and CL in arange(0, 25, 0.5)
  knees          [5] the Curvatura index of every block size (the repaired loop's cl2 ... cl32)
"""
import os
import sys
import types
import warnings

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference/modules/aclahe/python"
GOLD = os.path.join(ROOT, "tests", "golden")


def main():
    import cv2

    assert cv2.useOptimized()
    plt = types.ModuleType("matplotlib.pyplot")
    for name in ("plot", "xlabel", "ylabel", "title", "show"):
        setattr(plt, name, lambda *a, **k: None)
    mpl = types.ModuleType("matplotlib")
    mpl.pyplot = plt
    sys.modules["matplotlib"] = mpl
    sys.modules["matplotlib.pyplot"] = plt
    sys.path.insert(0, REF)
    import ACLAHE as RA  # the reference's file, unmodified
    import functions as AF

    img = cv2.imread(os.path.join(REF, "crowd.png"), 0)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        as_committed = RA.ParametrosACLAHE(img)
        # the repaired loop: ACLAHE.py:15-47 with lines 42-45 inside `for i in cl`, then :69-129 verbatim in meaning
        imgfilt = cv2.GaussianBlur(img, (3, 3), 0)
        bl = [2, 4, 8, 16, 32]
        cl = np.arange(0, 25, 0.5)
        resultados = np.zeros((10, 51), np.float32)
        n, v = 0, 1
        for k in bl:
            m = 1
            for i in cl:
                resultados[n, m] = i
                resultados[v, m] = AF.Entropia(AF.CLAHE(imgfilt, k, i))
                m += 1
            n += 2
            v += 2
        knees = []
        for j in range(5):
            x, y = AF.graficar(resultados[2 * j], resultados[2 * j + 1], "r")
            x22, x222, y220, y221, y222 = AF.DerivadaY(y)
            y223, y224, y225 = AF.DerivadaX(x, x22, x222)
            knees.append(int(AF.Curvatura(y220, y221, y222, y223, y224, y225)))
        d = max(knees)
        res2 = np.zeros((2, 5), np.float16)
        for m, k in enumerate(bl):
            res2[0, m] = k
            res2[1, m] = AF.Entropia(AF.CLAHE(imgfilt, k, d))
        w = [i for i, item in enumerate(res2[1]) if item == max(res2[1])][-1]
        repaired = (int(round(float(res2[0, w]))), d)
    ent = np.stack([resultados[2 * j + 1, 1:51] for j in range(5)]).astype(np.float32)
    print("as committed", as_committed, "repaired", repaired, "knees", knees)
    np.savez_compressed(os.path.join(GOLD, "crowd_full.npz"), img=img, as_committed=np.array(as_committed, np.int64),
                        repaired=np.array(repaired, np.int64), entropies=ent, knees=np.array(knees, np.int64),
                        pick_entropies=res2[1].astype(np.float16))


if __name__ == "__main__":
    main()
