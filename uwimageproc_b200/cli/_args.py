"""cv::CommandLineParser conventions (`-key=value`, `--key=value`, positional arguments) for the C++ CLI look-alikes."""


def parse(argv, keys):
    """-> (positional list, options dict, errors list).  `keys` maps option names to (default, type)."""
    pos, opt, err = [], {k: v[0] for k, v in keys.items()}, []
    for a in argv:
        if a.startswith("-") and len(a) > 1 and not a[1:2].isdigit():
            body = a.lstrip("-")
            name, sep, val = body.partition("=")
            if name in ("help", "h", "usage", "?"):
                opt["help"] = True
                continue
            if name not in keys:
                continue  # cv::CommandLineParser ignores unknown keys
            if not sep:
                val = "true"
            try:
                opt[name] = keys[name][1](val)
            except ValueError:
                err.append("Parameter '%s': can not convert '%s'" % (name, val))
        else:
            pos.append(a)
    return pos, opt, err
