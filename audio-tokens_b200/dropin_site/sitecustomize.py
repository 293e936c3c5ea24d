"""Put this directory on PYTHONPATH and the reference's own command line runs the B200 stages:

    PYTHONPATH=<repo>/audio-tokens_b200/dropin_site python /path/to/audio-tokens/run_pipeline.py

Python imports ``sitecustomize`` at start-up, before the script; at_b200.dropin.install() then routes exactly
``processors.{spectrogram_generator,cluster_creator,spec_tokenizer}`` to audio-tokens_b200/processors/ and leaves every
other import (audio_tokens_config, processors.model_trainer, datasets, models, utils) to the reference checkout, whose
directory Python itself puts at sys.path[0].  AT_B200_OPERATORS=1 additionally routes ``faiss`` (when not installed) and
torchaudio.transforms.MelSpectrogram / AmplitudeToDB, AT_B200_STAGES=0 leaves the reference's stage files in place
(then they run unchanged over the routed operators).
"""
import os
import sys

_pkg_root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if _pkg_root not in sys.path:
    sys.path.append(_pkg_root)

from at_b200 import dropin as _dropin  # noqa: E402

_dropin.install(operators=os.environ.get("AT_B200_OPERATORS", "0") not in ("", "0"),
                stages=os.environ.get("AT_B200_STAGES", "1") not in ("", "0"))
