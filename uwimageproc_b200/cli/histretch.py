"""histretch <input> <output> -c=<channels> [-cuda=0|1] [-time=0|1]      (modules/histretch/src/histretch.cpp:61-271)

Same positional arguments, flag strings and console messages as the reference binary; the HighGUI windows
(imshow / waitKey) are not opened.  Differences a maintainer should know about:
  * every run uses the GPU (this build has no CPU path); `-cuda=0` is accepted and reported;
  * the channel loop runs in the intended order (convert -> stretch -> merge -> convert back,
    modules/histretch/README.md:4); `-literal=1` reproduces the bytes the reference writes as committed, where the
    back-conversion runs before the merge (histretch.cpp:238-240) and the stretch of non-BGR letters is lost.
"""
import sys
import time

ABOUT = "histretch - percentile based histogram stretch of selected colour channels (B200 build)"


def main(argv=None):
    argv = sys.argv[1:] if argv is None else argv
    from ._args import parse

    pos, opt, err = parse(argv, {"c": ("r", str), "cuda": (1, int), "time": (0, int), "literal": (0, int), "help": (False, bool)})
    print(ABOUT)
    if len(pos) < 2 or opt.get("help"):
        print("C++ implementation of Histogram Stretching for specific channels of input image")
        print("Argument 'c=<channels>' is a string containing an ordered list of desired channels to be stretched")
        print("Histogram stretching is applied one at time, and then converted back to RGB colour space")
        print("Complete options are:")
        for line in ("-c=R|G|B\tfor RGB space", "-c=H|S|V\tfor HSV space", "-c=h|s|l\tfor HSL space", "-c=L|a|b\tfor Lab space",
                     "-c=Y|C|X\tfor YCrCb space", "-cuda=0 or -cuda=1 (CUDA ON: 1, CUDA OFF: 0, if available)"):
            print("\t" + line)
        print("\n\tExample:\n\t$ histretch -c=HV input.jpg output.jpg -cuda=0 -time=1")
        return 0
    if err:
        for e in err:
            print(e)
        return -1
    import cv2

    from ..api import default_context
    from ..modules import preprocessing as P

    if opt["cuda"] == 0:
        print("CUDA deactivated: this build has no CPU path, running on the GPU")
    print("***************************************")
    print("Input:", pos[0])
    print("Output:", pos[1])
    print("Channel:", opt["c"])
    src = cv2.imread(pos[0], cv2.IMREAD_COLOR)
    if src is None:
        print("Failed to read input image, exiting...")
        return -1
    ctx = default_context()
    print("Applying %d histretch" % len(opt["c"]))
    t0 = time.perf_counter()
    for nc, c in enumerate(opt["c"]):
        print("\tChannel[%d]: %s" % (nc, c))
        if P.numSpace(c) == -1:
            print("Option %s not recognized, skipping..." % c)
    out = ctx.histretch(src, opt["c"], 2, 98, "literal" if opt["literal"] else "intended")   # 2 / 98: histretch.cpp:154
    if opt["time"] == 1:
        print("\nExecution Time GPU :%g ms " % (1000.0 * (time.perf_counter() - t0)))
    print("hS: saving to disk")
    cv2.imwrite(pos[1], out)
    return 0


if __name__ == "__main__":
    sys.exit(main())
