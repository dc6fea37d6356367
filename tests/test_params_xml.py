"""parse_config_xml for the live settings.xml schema (UI/MdiEditor.cpp:566-749 reader, 751-1040 writer). Host-only code of
libvmorph.so: runs without a GPU."""
import os

import pytest

import videomorphing_b200 as vm
from videomorphing_b200 import api


def _write(tmp_path, lp, rp, cnt, weight='ssim="100.000000" tps="0.050000" ui="100000.000000" temp="10.000000" ssimclamp="0.000000"',
           debug='iternum="800" dropfactor="2.000000" eps="0.010000" startres="8"', lock=2, carry=False):
    def pts(tracks, terminate_last=True):
        s = ""
        for k, t in enumerate(tracks):
            for (x, y, z, w, wt) in t:
                s += "%d %d %d %d %f " % (x, y, z, w, wt)
            if terminate_last or k + 1 < len(tracks):
                s += "%d %d %d %d %f " % (-1, -1, -1, -1, -1.0)       # WriteXmlFile closes every track with an all -1 tuple
        return s
    con = ""
    for g in cnt:
        for c in g:
            con += "%d %d %d %d " % c
        con += "-1 -1 -1 -1 "
    xml = ("<?xml version='1.0'?>\n<project>\n <stage stage=\"3\"/>\n"
           " <videos video1=\"\\video1.mp4\" video2=\"\\video2.mp4\" resample1=\"\\resample1.mp4\" resample2=\"\\resample2.mp4\"/>\n"
           " <parameters>\n  <weight %s/>\n  <points image1=\"%s\" image2=\"%s\" connection=\"%s\" num=\"4\"/>\n"
           "  <boundary lock=\"%d\"/>\n  <debug %s/>\n </parameters>\n</project>\n"
           % (weight, pts(lp, not carry), pts(rp), con, lock, debug))
    p = tmp_path / "settings.xml"
    p.write_text(xml)
    return str(p)


def test_roundtrip_of_the_writer_format(tmp_path):
    lp = [[(10, 20, 0, 1, 1.0), (11, 21, 1, 0, 0.5)], [(40, 50, 0, 1, 0.25)]]
    rp = [[(12, 22, 0, 1, 1.0), (13, 23, 1, 0, 0.75)], [(44, 55, 0, 1, 1.0)]]
    cnt = [[(0, 0, 0, 0), (0, 1, 0, 1)], [(1, 0, 1, 0)]]
    prm = api.parse_config_xml(_write(tmp_path, lp, rp, cnt))
    assert (prm.w_ssim, prm.w_tps, prm.w_ui, prm.w_temp, prm.ssim_clamp) == (100.0, pytest.approx(0.05), 100000.0, 10.0, 0.0)
    assert (prm.max_iter, prm.max_iter_drop_factor, prm.eps, prm.start_res, prm.bcond) == (800, 2.0, pytest.approx(0.01), 8, vm.BCOND_BORDER)
    assert prm.lp == lp and prm.rp == rp and prm.cnt == cnt


def test_reader_quirks(tmp_path):
    # absent attributes read as 0 (QString::toFloat of an empty string), absent elements keep the defaults,
    # and an unterminated last image1 track carries over into the first image2 track (shared pt_list, MdiEditor.cpp:657-695)
    lp = [[(1, 2, 0, 1, 1.0)], [(7, 8, 0, 1, 1.0)]]
    rp = [[(3, 4, 0, 1, 1.0)]]
    prm = api.parse_config_xml(_write(tmp_path, lp, rp, [], weight='ssim="5" tps="0.5"', lock=1, carry=True))
    assert (prm.w_ssim, prm.w_tps, prm.w_ui, prm.w_temp) == (5.0, 0.5, 0.0, 0.0)
    assert prm.bcond == vm.BCOND_CORNER
    assert prm.lp == [lp[0]]
    assert prm.rp == [[lp[1][0], rp[0][0]]]
    assert prm.cnt == []
    p = tmp_path / "min.xml"
    p.write_text("<project><parameters><boundary lock='0'/></parameters></project>")
    d = vm.Parameters()
    q = api.parse_config_xml(str(p))
    assert (q.w_ssim, q.w_ui, q.max_iter, q.start_res, q.bcond) == (d.w_ssim, d.w_ui, d.max_iter, d.start_res, vm.BCOND_NONE)


def test_errors(tmp_path):
    with pytest.raises(vm._lib.VmError):
        api.parse_config_xml(str(tmp_path / "missing.xml"))          # parse_config_xml throws std::runtime_error (param_io.h)
    p = tmp_path / "bad.xml"
    p.write_text("<notaproject/>")
    with pytest.raises(vm._lib.VmError):
        api.parse_config_xml(str(p))


def test_writer_produces_the_reference_format_and_round_trips(tmp_path):
    """vm_params_write_xml = the settings.xml part of MdiEditor::WriteXmlFile (UI/MdiEditor.cpp:751-1040): the attribute values
    are the token strings the reference's sprintf calls produce (checked against the helper above, which follows them), num
    counts the key points of both images, and reader(writer(x)) == x."""
    lp = [[(10, 20, 0, 1, 1.0), (11, 21, 1, 0, 0.5)], [(40, 50, 0, 1, 0.25)]]
    rp = [[(12, 22, 0, 1, 1.0), (13, 23, 1, 0, 0.75)], [(44, 55, 0, 1, 1.0)]]
    cnt = [[(0, 0, 0, 0), (0, 1, 0, 1)], [(1, 0, 1, 0)]]
    prm = vm.Parameters(w_ssim=100.0, w_tps=0.05, w_ui=100000.0, w_temp=10.0, ssim_clamp=0.0, max_iter=800, max_iter_drop_factor=2.0, eps=0.01,
                        start_res=8, bcond=vm.BCOND_BORDER)
    prm.lp, prm.rp, prm.cnt = lp, rp, cnt
    out = tmp_path / "written.xml"
    api.write_config_xml(str(out), prm, stage=3)
    text = out.read_text()
    want = open(_write(tmp_path, lp, rp, cnt)).read()
    import re
    attrs = lambda s: dict(re.findall(r'(\w+)="([^"]*)"', s))
    a, b = attrs(text), attrs(want)
    assert a.pop("num") == "4" and b.pop("num") == "4"               # key points (p.w == 1) of image1 + image2
    assert a == b                                                     # every attribute value, token for token
    assert text.startswith("<?xml version='1.0'?>\n<project>\n    <stage stage=\"3\"/>")
    back = api.parse_config_xml(str(out))
    assert back.lp == lp and back.rp == rp and back.cnt == cnt
    for k in ("w_ssim", "w_tps", "w_ui", "w_temp", "ssim_clamp", "max_iter", "max_iter_drop_factor", "eps", "start_res", "bcond"):
        assert getattr(back, k) == pytest.approx(getattr(prm, k)), k
    # no points at all, corner lock
    q = vm.Parameters(bcond=vm.BCOND_CORNER)
    api.write_config_xml(str(tmp_path / "empty.xml"), q, stage=5)
    r = api.parse_config_xml(str(tmp_path / "empty.xml"))
    assert r.lp == [] and r.rp == [] and r.cnt == [] and r.bcond == vm.BCOND_CORNER and r.max_iter == q.max_iter
    with pytest.raises(vm._lib.VmError):
        api.write_config_xml(str(tmp_path / "no_such_dir" / "x.xml"), q)
