"""aclahe <input> <output>      (modules/aclahe/src/aclahe.cpp:64-226, modules/aclahe/python/main.py:17-22)

The reference binary converts to HSV, sweeps CLAHE over 5 block sizes x 51 clip limits on the V channel, prints the
entropy table and stops at its TODO list (find the knee, pick the block size, apply, convert back, save:
aclahe.cpp:209-218).  This shim prints the same table (one pixel pass per grid on the GPU) and then carries the TODO list
out with the procedure of the Python prototype (ACLAHE.py:69-129): knee of the entropy curve -> clip limit, float16
arg-max -> block size, CLAHE on V, HSV -> BGR, imwrite.  `-grey=1` is the Python main.py: grey image in, CLAHE image out.
"""
import sys


def main(argv=None):
    argv = sys.argv[1:] if argv is None else argv
    from ._args import parse

    pos, opt, err = parse(argv, {"grey": (0, int), "loop": ("repaired", str), "help": (False, bool)})
    print("ACLAHE: Automatic Contrast Limited Adaptive Histogram Equalization (B200 build)")
    if len(pos) < 2 or opt.get("help"):
        print("\n\tExample:\n\t$ aclahe input.jpg output.jpg")
        print("\tThis will apply ACLAHE to gray levels of 'input.jpg' image file, and save it into 'output.jpg'\n")
        return 0
    if err:
        for e in err:
            print(e)
        return -1
    import cv2
    import numpy as np

    from ..api import default_context
    from ..modules import aclahe as A

    print("***************************************")
    print("Input:", pos[0])
    print("Output:", pos[1])
    ctx = default_context()
    if opt["grey"]:
        img = cv2.imread(pos[0], 0)
        if img is None:
            print("Failed to read input image, exiting...")
            return -1
        BS, CL = A.ParametrosACLAHE(img, loop=opt["loop"])
        print("BS = %d, CL = %g" % (BS, CL))
        cv2.imwrite(pos[1], A.CLAHE(img, BS, CL))
        return 0
    src = cv2.imread(pos[0], cv2.IMREAD_COLOR)
    if src is None:
        print("Failed to read input image, exiting...")
        return -1
    print("Input image loaded...")
    hsv = cv2.cvtColor(src, cv2.COLOR_BGR2HSV)   # file-level glue only; the frame wrapper below converts on the GPU
    v = np.ascontiguousarray(hsv[..., 2])
    clips = np.arange(0.0, 25.0 + 1e-9, 0.5)     # aclahe.cpp:160-163: 0 ... 25 inclusive, 51 values
    for bs in A.BLOCK_SIZES:                      # the table the reference prints (aclahe.cpp:199-206), C++ entropy flavour
        print(" ".join("%g" % e for e in ctx.clahe_entropy_sweep(v, bs, clips, "cpp")))
    BS, CL = A.ParametrosACLAHE(v, loop=opt["loop"])
    print("BS = %d, CL = %g" % (BS, CL))
    cv2.imwrite(pos[1], ctx.aclahe(src, float(CL), (BS, BS)))
    return 0


if __name__ == "__main__":
    sys.exit(main())
