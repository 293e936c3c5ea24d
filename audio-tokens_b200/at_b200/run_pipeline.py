"""Launcher: runs the reference's unmodified ``run_pipeline.py`` with the three hot-path stages routed to the B200
implementation (at_b200.dropin).

    python -m at_b200.run_pipeline --reference /path/to/audio-tokens [--operators] [--script run_pipeline.py]

The script is executed with runpy as ``__main__`` and with its own directory at sys.path[0], exactly as
``python $REF/run_pipeline.py`` would (run_pipeline.py:1-18 of danavery/audio-tokens).
"""
import argparse
import os
import runpy
import sys

from . import dropin


def main(argv=None):
    ap = argparse.ArgumentParser(prog="python -m at_b200.run_pipeline")
    ap.add_argument("--reference", default=os.environ.get("AUDIO_TOKENS_REFERENCE"),
                    help="checkout of danavery/audio-tokens (default: $AUDIO_TOKENS_REFERENCE)")
    ap.add_argument("--script", default="run_pipeline.py", help="script inside the checkout to run")
    ap.add_argument("--operators", action="store_true",
                    help="also route faiss / torchaudio.transforms operators (level L-B)")
    ap.add_argument("--keep-reference-stages", action="store_true",
                    help="run the reference's own stage files over the routed operators (implies --operators)")
    args = ap.parse_args(argv)
    if not args.reference or not os.path.isdir(args.reference):
        ap.error("--reference (or $AUDIO_TOKENS_REFERENCE) must name the audio-tokens checkout")
    ref = os.path.abspath(args.reference)
    script = os.path.join(ref, args.script)
    if not os.path.isfile(script):
        ap.error(f"{script} does not exist")
    if sys.path and sys.path[0] != ref:
        sys.path.insert(0, ref)   # what ``python $REF/run_pipeline.py`` does
    dropin.install(reference=ref, operators=args.operators or args.keep_reference_stages,
                   stages=not args.keep_reference_stages)
    sys.argv = [script]
    runpy.run_path(script, run_name="__main__")


if __name__ == "__main__":
    main()
