"""compat.bind_reference: the reference's own run() (contours, deskew, line separation, reading order,
PAGE-XML -- its code, executed from /root/reference) on top of the drop-in hot-path methods must write
the same PAGE-XML content as the unmodified reference (tests/golden/ref_pipeline_run.xml, minted by
make_golden_pipeline_xml.py).  /root/reference only exists in the build container: skipped elsewhere."""
import os
import sys
import warnings

import cv2
import pytest

from conftest import GOLDEN
from sbb_textline_detection_b200 import compat, synth

sys.path.insert(0, GOLDEN)
import ref_import  # noqa: E402
import semantic_fake  # noqa: E402
from make_golden_pipeline_xml import PAGE  # noqa: E402


@pytest.mark.skipif(not os.path.exists(ref_import.REF_MAIN), reason="reference tree not present on this machine")
def test_bound_reference_run_writes_the_reference_xml(tmp_path):
    warnings.filterwarnings("ignore")
    ref_import.install_glue_stubs()
    ref = compat.import_reference(ref_import.REF_MAIN, name="_sbb_reference_for_binding")
    # hot path through the drop-in methods (duck-typed stand-in models -> the generic loop); the deskew
    # search stays on the reference's CPU code here because this machine has no GPU
    cls = compat.bind_reference(ref, gpu_deskew=False, model_loader=semantic_fake.loader)
    assert issubclass(cls, ref.textline_detector)
    png = str(tmp_path / "page.png")
    cv2.imwrite(png, synth.document_page(*PAGE[:2], seed=PAGE[2]))
    det = cls(png, str(tmp_path), "page", str(tmp_path))
    det.run()
    got = semantic_fake.summarise_xml(open(str(tmp_path / "page.xml")).read())
    want = semantic_fake.summarise_xml(open(os.path.join(GOLDEN, "ref_pipeline_run.xml")).read())
    assert got[0] == want[0]
    assert len(got[1]) == len(want[1]) == 6          # identical PAGE-XML region count (BASELINE north_star)
    assert got[1] == want[1]                         # ... and identical region / line polygons


def test_xml_summary_helper():
    xml = open(os.path.join(GOLDEN, "ref_page_full.xml")).read()
    border, regions = semantic_fake.summarise_xml(xml)
    assert border.startswith("10,18") and len(regions) == 4 and sum(len(r[1]) for r in regions) == 11


def test_cli_help_and_options():
    """The reference's only CI check is ``sbb_textline_detector --help`` (.travis.yml:15-16)."""
    from click.testing import CliRunner
    from sbb_textline_detection_b200 import cli
    res = CliRunner().invoke(cli.main, ["--help"])
    assert res.exit_code == 0
    for opt in ("--image", "-i", "--out", "-o", "--model", "-m"):   # main.py:2160-2166
        assert opt in res.output
    res = CliRunner().invoke(cli.main, [])
    assert res.exit_code != 0 and "Missing option" in res.output
