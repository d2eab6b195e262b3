"""Shim for the reference's ``models`` package (reference models/__init__.py is empty).

Only the two hot-path modules are replaced here: ``models.decoderlstm`` (models/decoderlstm.py:11 AttentionGru, plus the
DecoderGRU / DecoderRNN that hypernet.py:11 imports from it) and ``models.attention`` (models/attention.py:5).  Every
other submodule the reference scripts import -- ``models.encoder`` (cc_train_hypernet.py:12, train_hyper_combine.py:12,
test_hn.py:16, hypernet_attention.py:13), ``models.layers`` -- must keep resolving to the reference's own files, so this
package's ``__path__`` is extended with every other ``models`` directory found on ``sys.path`` (the reference checkout).
The shim directory stays first: its decoderlstm / attention win, anything it does not define falls through.
"""
import pkgutil

__path__ = pkgutil.extend_path(__path__, __name__)
