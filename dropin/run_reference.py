#!/usr/bin/env python
"""Run an UNEDITED reference script on top of the B200 drop-in classes.

    python /path/to/dropin/run_reference.py [--reference DIR] cc_train_hypernet.py [script args ...]

``python script.py`` always puts the script's own directory first on ``sys.path``, so the reference's
``hypernet_attention.py`` / ``models/decoderlstm.py`` would win over anything on PYTHONPATH.  This launcher sets
``sys.path = [dropin/, <reference dir>, ...]`` (the order dropin/README.md asks for), makes the reference directory the
working directory (the scripts open relative ``data/...`` paths, e.g. later.py:372) and executes the script as
``__main__`` with its own argv.  Nothing of the reference is edited or copied.
"""
import os
import runpy
import sys

HERE = os.path.dirname(os.path.abspath(__file__))


def install(reference_dir=None):
    """Put dropin/ (and the package root) ahead of ``reference_dir`` on sys.path."""
    root = os.path.dirname(HERE)
    for p in (HERE, root):
        while p in sys.path:
            sys.path.remove(p)
    if reference_dir is not None:
        reference_dir = os.path.abspath(reference_dir)
        while reference_dir in sys.path:
            sys.path.remove(reference_dir)
        sys.path.insert(0, reference_dir)
    sys.path.insert(0, root)
    sys.path.insert(0, HERE)


def main(argv):
    ref = None
    if len(argv) >= 2 and argv[0] == "--reference":
        ref, argv = argv[1], argv[2:]
    if not argv:
        sys.exit(__doc__)
    script = os.path.abspath(os.path.join(ref, argv[0]) if ref and not os.path.isabs(argv[0]) else argv[0])
    ref = ref or os.path.dirname(script)
    install(ref)
    os.chdir(ref)
    sys.argv = [script] + argv[1:]
    runpy.run_path(script, run_name="__main__")


if __name__ == "__main__":
    main(sys.argv[1:])
