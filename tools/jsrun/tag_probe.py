"""Runs the parts of lamejs behind SURVEY.md 8(f3) under the engine (see ref_lamejs.py): test infrastructure.

1. `hot_path_crc`: the UNMODIFIED `lamejs.Mp3Encoder`; one of ref_lamejs's in-memory hooks remembers `gfp`, so the state
   copy_buffer keeps on every call -- gfc.nMusicCRC, VBR_seek_table.nBytesWritten (BitStream.js:924-935) -- can be read
   after flush().
2. `tagged`: the same module wiring as index.js:66-115 with gfp.bWriteVbrTag = true.  src/js/VBRTag.js is not runnable as
   shipped (it was never executed: Mp3Encoder switches the tag off); the names it leaves unbound are bound here, WITHOUT
   touching the source text:
       Tables, Lame (+ Lame.LAME_ID), vbr_off / vbr_abr, MONO / STEREO / ..., `int` (so that `new int[400]` yields an
       Int32Array(400)), lame.BitrateIndex (a private function of Lame.js:431-443 the tag writer calls as a method),
       VbrMode.vbr_off.ordinal as a function (common.js:138-141 makes it a number; VBRTag.js:637 calls it).
   What remains JavaScript-only behaviour is reported as is: TotalFrameSize is a fractional number when the frame size does
   not divide (44.1 / 22.05 / 11.025 kHz), `0xff & "L"` is 0 (magic and version strings come out as zero bytes),
   `bag[i / 2]` does not compact the seek bag.  oracle/lj_vbrtag.cpp follows the Java original there and says so.
3. `wav_header`: lamejs.WavHeader.readHeader on a byte string.
"""
import json
import os
import tempfile

import ref_lamejs as R

_CAPTURE_JS = "var __gfp = null; function __lj_pre(gfp, masking) { __gfp = gfp; }\n"

_CRC_DRIVER = R._DRIVER.replace(
    "return JSON.stringify({sizes:sizes,",
    "var gfc = __gfp ? __gfp.internal_flags : null;\n"
    "  return JSON.stringify({crc: gfc ? gfc.nMusicCRC : -1, nbytes: gfc ? gfc.VBR_seek_table.nBytesWritten : -1, sizes:sizes,")


def hot_path_crc(channels, samplerate, kbps, left, right=None, chunk=None):
    """(bytes, nMusicCRC, nBytesWritten) of an unmodified Mp3Encoder run."""
    data, sizes, o = R.encode(channels, samplerate, kbps, left, right, chunk=chunk, loader="taps", extra_js=_CAPTURE_JS, driver=_CRC_DRIVER)
    return data, o["crc"], o["nbytes"]


_TAG_DRIVER = r"""
var Lame=__require('Lame.js'), Presets=__require('Presets.js'), GainAnalysis=__require('GainAnalysis.js'), QuantizePVT=__require('QuantizePVT.js'),
 Quantize=__require('Quantize.js'), Takehiro=__require('Takehiro.js'), Reservoir=__require('Reservoir.js'), MPEGMode=__require('MPEGMode.js'),
 BitStream=__require('BitStream.js'), Version=__require('Version.js'), VBRTag=__require('VBRTag.js'), common=__require('common.js');
var Tables=__require('Tables.js');
var VbrMode=common.VbrMode, vbr_off=VbrMode.vbr_off, vbr_abr=VbrMode.vbr_abr;
var MONO=MPEGMode.MONO, STEREO=MPEGMode.STEREO, DUAL_CHANNEL=MPEGMode.DUAL_CHANNEL, JOINT_STEREO=MPEGMode.JOINT_STEREO, NOT_SET=MPEGMode.NOT_SET;
var int = []; int[400] = function(){ return new Int32Array(400); };
function __Stub(){ this.setModules=function(){}; }
function __unhex(h){ var n=h.length/4; var a=new Int16Array(n); for(var i=0;i<n;i++){
  var v=parseInt(h.substr(4*i+2,2)+h.substr(4*i,2),16); a[i]= v>=32768? v-65536: v;} return a;}
function __tohex(b,n){ var s=[]; for(var i=0;i<n;i++){ var v=b[i]&255; s.push((v<16?"0":"")+v.toString(16)); } return s.join(""); }
(function(){
 var lame=new Lame(), ga=new GainAnalysis(), bs=new BitStream(), p=new Presets(), qupvt=new QuantizePVT(), qu=new Quantize(), vbr=new VBRTag(),
  ver=new Version(), id3=new __Stub(), rv=new Reservoir(), tak=new Takehiro(), mpg=new __Stub();
 lame.setModules(ga,bs,p,qupvt,qu,vbr,ver,id3,mpg); bs.setModules(ga,mpg,ver,vbr); p.setModules(lame); qu.setModules(bs,rv,qupvt,tak);
 qupvt.setModules(tak,rv,lame.enc.psy); rv.setModules(bs); tak.setModules(qupvt); vbr.setModules(lame,bs,ver);
 lame.BitrateIndex=function(b,v,sr){ if(sr<16000) v=2; for(var i=0;i<=14;i++){ if(Tables.bitrate_table[v][i]>0 && Tables.bitrate_table[v][i]==b) return i;} return -1; };
 var gfp=lame.lame_init();
 gfp.num_channels=__CH; gfp.in_samplerate=__SR; gfp.brate=__KBPS; gfp.mode=MPEGMode.STEREO; gfp.quality=3;
 gfp.bWriteVbrTag=true; gfp.disable_reservoir=true; gfp.write_id3tag_automatic=false;
 var rc=lame.lame_init_params(gfp);
 var gfc=gfp.internal_flags;
 Lame.LAME_ID = gfc.Class_ID;
 var L=__unhex(__HEXL), R=__CH==2?__unhex(__HEXR):L, n=L.length, step=__CHUNK>0?__CHUNK:Math.max(n,1);
 var hex=[], sizes=[];
 for (var i=0;i<n;i+=step){ var l=L.subarray(i,i+step), r=R.subarray(i,i+step);
   var buf=new Int8Array(0|(1.25*l.length+7200+2880)); var k=lame.lame_encode_buffer(gfp,l,r,l.length,buf,0,buf.length);
   sizes.push(k); hex.push(__tohex(buf,k)); }
 var fb=new Int8Array(7200+4*2880); var k2=lame.lame_encode_flush(gfp,fb,0,fb.length); sizes.push(k2); hex.push(__tohex(fb,k2));
 var tag=new Int8Array(2880), ts;
 var __ord=function(){ return 0; }; __ord.valueOf=function(){ return 0; }; VbrMode.vbr_off.ordinal=__ord;   /* `gfp.VBR.ordinal()` (VBRTag.js:637) */
 try { ts=vbr.getLameTagFrame(gfp,tag); } catch(e){ ts="ERR "+e; }
 var st=gfc.VBR_seek_table;
 var bag=[]; for (var i=0;i<st.pos;i++) bag.push(st.bag[i]);
 return JSON.stringify({rc:rc, sizes:sizes, hex:hex.join(""), tag_ret:ts, tag:__tohex(tag, 2880), crc:gfc.nMusicCRC, nbytes:st.nBytesWritten,
   frames:st.nVbrNumFrames, pos:st.pos, sum:st.sum, want:st.want, bag:bag, total_frame_size:st.TotalFrameSize, write_tag:gfp.bWriteVbrTag,
   encoder_padding:gfp.encoder_padding, lowpassfreq:gfp.lowpassfreq, noise_shaping:gfc.noise_shaping, preset:gfp.preset, vbr_q:gfp.VBR_q,
   exp_nspsytune:gfp.exp_nspsytune, athtype:gfp.ATHtype, sideinfo_len:gfc.sideinfo_len, mode_ext:gfc.mode_ext});
})();
"""


def tagged(channels, samplerate, kbps, left, right=None, chunk=None):
    """lamejs with gfp.bWriteVbrTag = true (names bound as described above).  Returns the engine's JSON as a dict with
    `bytes` (all calls concatenated) and `tag` (2880-byte buffer handed to getLameTagFrame) as bytes."""
    if right is None:
        right = left
    with tempfile.TemporaryDirectory() as td:
        p = os.path.join(td, "modules.js")
        open(p, "w").write(R.modules_loader_source(hooks=False))
        d = os.path.join(td, "drive.js")
        with open(d, "w") as f:
            f.write('var __HEXL="%s"; var __HEXR="%s"; var __CH=%d, __SR=%d, __KBPS=%d, __CHUNK=%d;\n'
                    % (R._hex16(left), R._hex16(right), channels, samplerate, kbps, chunk or 0))
            f.write(_TAG_DRIVER)
        o = json.loads(R.run_js([p, d]))
    o["bytes"] = bytes.fromhex(o.pop("hex"))
    o["tag"] = bytes.fromhex(o["tag"])
    return o


_WAV_DRIVER = r"""
(function(){
  var h=__HEX, n=h.length/2, ab=new ArrayBuffer(n), u=new Uint8Array(ab);
  for (var i=0;i<n;i++) u[i]=parseInt(h.substr(2*i,2),16);
  var w, err=null;
  try { w = lamejs.WavHeader.readHeader(new DataView(ab)); } catch(e) { err = (typeof e === 'string') ? e : (e && e.name ? e.name : String(e)); }
  if (err !== null) return JSON.stringify({"throws": err});
  if (w === undefined) return JSON.stringify({"undefined": true});
  return JSON.stringify({dataOffset:w.dataOffset, dataLen:w.dataLen, channels:w.channels, sampleRate:w.sampleRate});
})();
"""


def wav_header(data, loader="bundle"):
    with tempfile.TemporaryDirectory() as td:
        files = [os.path.join(R.REF, "lame.all.js")]
        d = os.path.join(td, "drive.js")
        open(d, "w").write('var __HEX="%s";\n' % bytes(data).hex() + _WAV_DRIVER)
        files.append(d)
        return json.loads(R.run_js(files))
