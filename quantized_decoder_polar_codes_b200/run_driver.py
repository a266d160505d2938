"""Run one of the reference's own scripts, unmodified, on this package:

    cd /path/to/Quantized_Decoder_Polar_Codes        # the scripts use ./reliable sequence.txt and ./LUT/...
    python -m quantized_decoder_polar_codes_b200.run_driver [--max-frames M] [--seed S] mainQuantizedDecoder_LLRDomain.py \\
           --N 1024 --A 512 --L 8 --DecoderType SCL-LUT

compat.install() supplies the import paths, numpy/torchtracer/matplotlib shims and the quantizer adapters the scripts need
(SURVEY.md Appendix C); everything after the script name is the script's own command line.  Returns the script's globals
when called as run(path, argv, ...)."""
import os
import runpy
import sys

from . import compat


def run(script, argv=(), max_frames=None, seed=None, **install_kw):
    compat.install(max_frames=max_frames, seed=seed, **install_kw)
    script = os.path.abspath(script)
    old_argv, old_path = sys.argv, list(sys.path)
    sys.argv = [script] + list(argv)
    sys.path.insert(0, os.path.dirname(script))     # what `python script.py` does
    try:
        return runpy.run_path(script, run_name="__main__")
    finally:
        sys.argv, sys.path[:] = old_argv, old_path


def main():
    args = sys.argv[1:]
    max_frames = seed = None
    while args and args[0] in ("--max-frames", "--seed"):
        if args[0] == "--max-frames":
            max_frames = int(args[1])
        else:
            seed = int(args[1])
        args = args[2:]
    if not args:
        raise SystemExit(__doc__)
    run(args[0], args[1:], max_frames=max_frames, seed=seed)


if __name__ == "__main__":
    main()
