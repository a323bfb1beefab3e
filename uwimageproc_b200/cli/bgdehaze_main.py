"""python -m uwimageproc_b200.cli.bgdehaze_main -i <index> -w <window>      (modules/bgdehaze/main.py:14-36)

The reference resolves `-i` through util.get_filenames() (a module that is not in the repository): a list of
(source, destination) pairs.  Here the list is built from `--dir` (default: ./img, the reference's fixture folder):
every image file in it, sorted, destination = <dir>/../results/<name>.  `--src/--dest` name one pair directly.
"""
import argparse
import os
import sys


def get_filenames(folder):
    exts = (".jpg", ".jpeg", ".png", ".bmp", ".tif", ".tiff")
    names = sorted(f for f in os.listdir(folder) if f.lower().endswith(exts)) if os.path.isdir(folder) else []
    out_dir = os.path.join(os.path.dirname(os.path.abspath(folder)), "results")
    return [(os.path.join(folder, f), os.path.join(out_dir, f)) for f in names]


def generate_results(src, dest, w=15):   # main.py:14-20
    import cv2

    from ..modules import bgdehaze as M

    print("processing", src + "...")
    I = cv2.imread(src)
    if I is None:
        raise SystemExit("cannot read " + src)
    restored8 = M.generate_results(I, w)   # normalisation, adaptiveExp_map and the imwrite rounding, on the GPU
    os.makedirs(os.path.dirname(os.path.abspath(dest)) or ".", exist_ok=True)
    cv2.imwrite(dest, restored8)
    print("saved", dest)


def main(argv=None):
    pre = argparse.ArgumentParser(add_help=False)
    pre.add_argument("--dir", default="img")
    known, _ = pre.parse_known_args(argv)
    filenames = get_filenames(known.dir)
    p = argparse.ArgumentParser(description="Underwater Image Restoration by Blue-Green Channels Dehazing and Red Channel Correction")
    p.add_argument("--dir", default="img", help="folder the -i index refers to")
    p.add_argument("-i", "--input", type=int, choices=range(len(filenames)) if filenames else None,
                   help="index for single input image" + (": {} corresponds to indexes {}".format(filenames[0][0], list(range(len(filenames)))) if filenames else ""))
    p.add_argument("-w", "--window", type=int, default=15, help="window size of dark channel")
    p.add_argument("--src")
    p.add_argument("--dest")
    args = p.parse_args(argv)
    if args.src:
        src, dest = args.src, args.dest or (os.path.splitext(args.src)[0] + "_dehazed.png")
    else:
        if args.input is None or not filenames:
            p.error("give -i <index> (with images under --dir) or --src/--dest")
        src, dest = filenames[args.input]
    generate_results(src, dest, args.window)
    return 0


if __name__ == "__main__":
    sys.exit(main())
