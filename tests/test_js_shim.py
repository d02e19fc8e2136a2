"""js/index.js (the ffi-napi binding a lamejs user would load) cannot run here -- no node, no ffi -- but Qt's JavaScript engine can
at least parse it and execute its control flow against a stubbed `ffi-napi` / `ref-napi`: the file is syntactically valid, exports
the lamejs surface, and calls the C-ABI entry points in the order and with the arity include/mp3b200.h declares."""
import json
import os
import re
import sys
import tempfile

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)

STUBS = r"""
var module = {exports: {}}, process = {env: {}};
var __calls = [];
function Buffer() {}
Buffer.alloc = function (n) { var b = new Uint8Array(n); b.readBigInt64LE = function () { return 44; }; b.readInt32LE = function () { return 2; };
  b.readUInt32LE = function () { return 44100; }; return b; };
Buffer.from = function (a, o, l) { return new Uint8Array(a, o, l); };
function require(name) {
  if (name === 'ffi-napi') return {Library: function (n, decl) { var lib = {__decl: decl}; Object.keys(decl).forEach(function (k) {
      lib[k] = function () { __calls.push([k, arguments.length]); if (k === 'mp3b200_lametag_size') return 417;
        if (k === 'mp3b200_wav_read_header' || k === 'mp3b200_set_write_vbr_tag') return 1; if (k === 'mp3b200_export_state') return 8; return 0; }; });
      __lib = lib; return lib; }};
  if (name === 'ref-napi') return {refType: function (t) { return {t: t}; }, types: {void: 'void'}, alloc: function () { return {deref: function () { return {}; }}; }, NULL: null};
  throw new Error('unexpected require ' + name);
}
var __lib = null;
"""

DRIVER = r"""
(function () {
  var m = module.exports;
  var e = new m.Mp3Encoder(2, 44100, 128, {writeVbrTag: true});
  var a = e.encodeBuffer(new Int16Array(1152), new Int16Array(1152));
  var f = e.flush();
  var t = e.getLameTagFrame();
  var s = e.exportState(); e.importState(s); e.seek(3, new Int16Array(1328), new Int16Array(1328));
  var w = m.WavHeader.readHeader(new DataView(new ArrayBuffer(64)));
  e.close();
  var decl = {}; Object.keys(__lib.__decl).forEach(function (k) { decl[k] = __lib.__decl[k][1].length; });
  return JSON.stringify({calls: __calls, decl: decl, exports: Object.keys(m), w: w, types: [a instanceof Int8Array, f instanceof Int8Array, t instanceof Int8Array]});
})();
"""


def test_shim_parses_and_calls_the_abi_in_order():
    sys.path.insert(0, os.path.join(ROOT, "tools", "jsrun"))
    import ref_lamejs
    if not ref_lamejs.qt_dir() or not os.path.exists(os.path.join(ref_lamejs.qt_dir(), "libQt6Qml.so.6")):
        pytest.skip("no JavaScript engine in this environment")
    shim = open(os.path.join(ROOT, "js", "index.js")).read()
    with tempfile.TemporaryDirectory() as td:
        p = os.path.join(td, "shim.js")
        open(p, "w").write(STUBS + shim + DRIVER)
        o = json.loads(ref_lamejs.run_js([p]))
    assert o["exports"] == ["Mp3Encoder", "WavHeader"] and o["types"] == [True, True, True]
    assert [c[0] for c in o["calls"]] == ["mp3b200_create", "mp3b200_set_write_vbr_tag", "mp3b200_lametag_size", "mp3b200_encode", "mp3b200_flush",
                                          "mp3b200_get_lametag_frame", "mp3b200_export_state", "mp3b200_export_state", "mp3b200_import_state",
                                          "mp3b200_seek", "mp3b200_wav_read_header", "mp3b200_destroy"]
    assert o["w"] == {"dataOffset": 44, "dataLen": 44, "channels": 2, "sampleRate": 44100}
    # every bound function exists in the header with that many parameters, and is called with that many arguments
    hdr = open(os.path.join(ROOT, "include", "mp3b200.h")).read()
    for name, nargs in o["decl"].items():
        m = re.search(r"\b%s\s*\(([^;]*?)\)\s*;" % name, hdr, re.S)
        assert m, name
        params = [x for x in m.group(1).split(",") if x.strip() and x.strip() != "void"]
        assert len(params) == nargs, (name, params, nargs)
    for name, n in o["calls"]:
        assert n == o["decl"][name], (name, n)
