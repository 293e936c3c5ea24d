"""Drop-in ``processors`` package: SpectrogramGenerator, ClusterCreator and SpecTokenizer re-implemented over
libat_b200.  Put ``audio-tokens_b200/`` ahead of the reference checkout on sys.path and the reference's
``run_pipeline.py`` imports these three classes unchanged; every other ``processors.*`` module it imports
(model_trainer, dataset_splitter, audioset_metadata_processor) is found in the reference checkout through the
package path extension below (set AUDIO_TOKENS_REFERENCE=/path/to/audio-tokens).
"""
import os

_ref = os.environ.get("AUDIO_TOKENS_REFERENCE")
if _ref and os.path.isdir(os.path.join(_ref, "processors")):
    __path__.append(os.path.join(_ref, "processors"))
