"""chain <input.jpg> <output.jpg> [-q=95]      the whole path on one file: histretch(V, 1/99) -> aclahe(8x8, clip 2) -> bgdehaze

JPEG files go through nvJPEG on the device (uwip_chain_jpeg): the file is read as bytes, decoded, enhanced and encoded on the
GPU, and written as bytes - the pixels never exist on the host.  Other formats take the cv2 route of the
per-module CLIs (imread -> uwip_chain_bgr8 -> imwrite).
"""
import sys


def main(argv=None):
    argv = sys.argv[1:] if argv is None else argv
    from ._args import parse

    pos, opt, err = parse(argv, {"q": (95, int), "help": (False, bool)})
    if len(pos) < 2 or opt.get("help") or err:
        print(__doc__)
        return 0 if not err else -1
    from ..api import default_context

    ctx = default_context()
    src, dst = pos[0], pos[1]
    if src.lower().endswith((".jpg", ".jpeg")) and dst.lower().endswith((".jpg", ".jpeg")):
        with open(src, "rb") as f:
            data = f.read()
        out = ctx.chain_jpeg(data, quality=opt["q"])
        with open(dst, "wb") as f:
            f.write(out)
    else:
        import cv2

        img = cv2.imread(src, cv2.IMREAD_COLOR)
        if img is None:
            print("Failed to read input image, exiting...")
            return -1
        cv2.imwrite(dst, ctx.chain(img))
    flags = ctx.last_frame_flags(1)
    if flags[0]:
        print("note: the reference's exposure map is NaN for this frame (S = 0/0); the output is all zeros, as imwrite would store it")
    print("saved", dst)
    return 0


if __name__ == "__main__":
    sys.exit(main())
