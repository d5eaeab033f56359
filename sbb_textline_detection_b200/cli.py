"""``sbb_textline_detector`` command line of the reference (``main``, main.py:2160-2171: -i image, -o out
dir, -m model dir) on the B200 hot path: the reference's own class and ``run()`` (contours, line
separation, reading order, PAGE-XML), bound to the GPU methods by compat.bind_reference.

    python -m sbb_textline_detection_b200.cli -i page.png -o out/ -m models/ [--reference /path/to/main.py]

``patch_reference_package()`` does the same for an installed ``qurator.sbb_textline_detector`` package,
so that its two console scripts (``sbb_textline_detector`` and the OCR-D processor
``ocrd-sbb-textline-detector``, ocrd_cli.py:84 instantiates ``textline_detector``) use the GPU path.
"""
from __future__ import annotations

import importlib
import sys

import click

from . import compat


def _reference_module(reference: str | None):
    if reference:
        return compat.import_reference(reference)
    try:
        return importlib.import_module("qurator.sbb_textline_detector.main")
    except Exception as e:  # tensorflow 1.15 / keras 2.3 pins do not import on a current stack
        spec = importlib.util.find_spec("qurator.sbb_textline_detector") if importlib.util.find_spec("qurator") else None
        if spec is None or not spec.submodule_search_locations:
            raise click.UsageError(
                "the reference package qurator.sbb_textline_detector is not importable; pass --reference "
                f"/path/to/qurator/sbb_textline_detector/main.py ({e})")
        import os
        return compat.import_reference(os.path.join(list(spec.submodule_search_locations)[0], "main.py"))


def patch_reference_package(**bind_kwargs):
    """Rebind ``textline_detector`` inside an installed reference package to the GPU-backed subclass."""
    ref = _reference_module(None)
    bound = compat.bind_reference(ref, **bind_kwargs)
    ref.textline_detector = bound
    pkg = sys.modules.get("qurator.sbb_textline_detector")
    if pkg is not None:
        pkg.textline_detector = bound
    return bound


@click.command()
@click.option('--image', '-i', help='image filename', type=click.Path(exists=True, dir_okay=False), required=True)
@click.option('--out', '-o', help='directory to write output xml data', type=click.Path(exists=True, file_okay=False),
              required=True)
@click.option('--model', '-m', help='directory of models', type=click.Path(exists=True, file_okay=False), required=True)
@click.option('--reference', help="path to the reference's main.py (default: the installed qurator.sbb_textline_detector)",
              type=click.Path(exists=True, dir_okay=False), default=None)
@click.option('--device', type=int, default=0, show_default=True, help='CUDA device')
def main(image, out, model, reference, device):
    cls = compat.bind_reference(_reference_module(reference), device=device)
    x = cls(image, out, None, model)
    x.run()


if __name__ == "__main__":
    main()
