"""Import hook that makes the reference's unmodified ``run_pipeline.py`` drive the B200 stages.

The reference imports its three hot-path stages by name (run_pipeline.py:1-5 of danavery/audio-tokens):

    from processors.cluster_creator import ClusterCreator
    from processors.spec_tokenizer import SpecTokenizer
    from processors.spectrogram_generator import SpectrogramGenerator

``python $REF/run_pipeline.py`` puts ``$REF`` at ``sys.path[0]``, so the package ``processors`` (and
``audio_tokens_config``, ``processors.model_trainer``, ``datasets``, ``models``, ``utils`` ...) always resolve to the
reference checkout, whatever PYTHONPATH says.  ``install()`` therefore does not play with the path order: it puts a
finder at the head of ``sys.meta_path`` that answers for exactly the three module names above with the files under
``audio-tokens_b200/processors/`` and for nothing else.  Everything else of the reference keeps resolving to the
reference, including its full ``AudioTokensConfig`` (the B200 stages read the fields they need by attribute and the
knobs that only exist here with ``getattr(config, name, default)``).

Optionally (``operators=True``) the library operators are routed too, which lets even the reference's OWN stage files run
on the B200 kernels unchanged (level L-B): ``faiss`` resolves to ``at_b200.faiss_compat`` when no
real faiss is installed, and ``torchaudio.transforms.MelSpectrogram`` / ``AmplitudeToDB`` are replaced by the
``nn.Module`` mirrors of ``at_b200.torchaudio_compat``.

Three ways in:
    python -m at_b200.run_pipeline [--reference DIR]            (launcher, audio-tokens_b200 on PYTHONPATH)
    PYTHONPATH=<repo>/audio-tokens_b200/dropin_site python $REF/run_pipeline.py      (sitecustomize: nothing else changes)
    import at_b200.dropin; at_b200.dropin.install()             (from your own driver)
"""
from __future__ import annotations

import importlib.abc
import importlib.util
import os
import sys

_PKG_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))   # .../audio-tokens_b200
STAGE_DIR = os.path.join(_PKG_ROOT, "processors")
STAGE_MODULES = {
    "processors.spectrogram_generator": os.path.join(STAGE_DIR, "spectrogram_generator.py"),
    "processors.cluster_creator": os.path.join(STAGE_DIR, "cluster_creator.py"),
    "processors.spec_tokenizer": os.path.join(STAGE_DIR, "spec_tokenizer.py"),
}


class _StageFinder(importlib.abc.MetaPathFinder):
    """Answers for the three stage modules only (and, when asked, for a missing ``faiss``)."""

    def __init__(self, stages: bool, faiss_shim: bool):
        self.stages = stages
        self.faiss_shim = faiss_shim

    def find_spec(self, fullname, path=None, target=None):
        file = STAGE_MODULES.get(fullname) if self.stages else None
        if file is not None:
            return importlib.util.spec_from_file_location(fullname, file)
        if self.faiss_shim and fullname == "faiss":
            # only when no real faiss can be found by the remaining finders
            for finder in sys.meta_path:
                if finder is self or not hasattr(finder, "find_spec"):
                    continue
                try:
                    if finder.find_spec(fullname, path, target) is not None:
                        return None
                except Exception:
                    continue
            return importlib.util.spec_from_loader("faiss", _AliasLoader("at_b200.faiss_compat"))
        return None


class _AliasLoader(importlib.abc.Loader):
    """``import faiss`` hands out the already-importable module at_b200.faiss_compat."""

    def __init__(self, target: str):
        self.target = target

    def create_module(self, spec):
        import importlib

        return importlib.import_module(self.target)

    def exec_module(self, module):
        pass


_installed = None


def installed() -> bool:
    return _installed is not None


def install(reference: str | None = None, operators: bool = False, stages: bool = True):
    """Route the reference's imports to the B200 implementation.  Idempotent.

    reference   the reference checkout; appended to sys.path when given (or $AUDIO_TOKENS_REFERENCE) and not there yet.
    stages      the three ``processors.*`` stage modules resolve to audio-tokens_b200/processors/ (level L-A).
    operators   ``import faiss`` -> at_b200.faiss_compat when faiss is not installed, and torchaudio.transforms'
                MelSpectrogram / AmplitudeToDB -> at_b200.torchaudio_compat (level L-B).
    """
    global _installed
    if _PKG_ROOT not in sys.path:
        sys.path.append(_PKG_ROOT)   # for ``import at_b200``; appended, so it never shadows the reference's packages
    reference = reference or os.environ.get("AUDIO_TOKENS_REFERENCE")
    if reference:
        reference = os.path.abspath(reference)
        os.environ.setdefault("AUDIO_TOKENS_REFERENCE", reference)
        if reference not in sys.path:
            sys.path.append(reference)
    if _installed is None:
        _installed = _StageFinder(stages=stages, faiss_shim=operators)
        sys.meta_path.insert(0, _installed)
    else:
        _installed.stages = _installed.stages or stages
        _installed.faiss_shim = _installed.faiss_shim or operators
    if stages:
        # modules imported before install() would keep pointing at the reference's files
        for name, file in STAGE_MODULES.items():
            if name in sys.modules and getattr(sys.modules[name], "__file__", None) != file:
                del sys.modules[name]
    if operators:
        _patch_torchaudio()
    return _installed


def _patch_torchaudio():
    """torchaudio.transforms.{MelSpectrogram, AmplitudeToDB} -> the B200 mirrors (reference call site:
    processors/spectrogram_generator.py:10,28-34)."""
    import torchaudio.transforms as T

    from . import torchaudio_compat as C

    if getattr(T, "_at_b200_original", None) is None:
        T._at_b200_original = (T.MelSpectrogram, T.AmplitudeToDB)
    T.MelSpectrogram = C.MelSpectrogram
    T.AmplitudeToDB = C.AmplitudeToDB


def uninstall():
    """Undo install() (tests)."""
    global _installed
    if _installed is not None:
        try:
            sys.meta_path.remove(_installed)
        except ValueError:
            pass
        _installed = None
    for name in STAGE_MODULES:
        mod = sys.modules.get(name)
        if mod is not None and getattr(mod, "__file__", None) == STAGE_MODULES[name]:
            del sys.modules[name]
    T = sys.modules.get("torchaudio.transforms")
    if T is not None and getattr(T, "_at_b200_original", None) is not None:
        T.MelSpectrogram, T.AmplitudeToDB = T._at_b200_original
        T._at_b200_original = None
    f = sys.modules.get("faiss")
    if f is not None and (getattr(f, "__file__", None) or "").endswith("faiss_compat.py"):
        del sys.modules["faiss"]
