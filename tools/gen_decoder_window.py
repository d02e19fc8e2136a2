#!/usr/bin/env python3
"""Dev-time generator (needs /root/reference): extracts the 257-entry synthesis prototype window of the MPEG audio
decoder that ships in the reference's Java tree (src/main/java/mpg/TabInit.java `dewin`, the ISO 11172-3 table D[i] up
to its centre, modulation signs stripped) into tests/golden/synth_window.json.  It is used by tests/mp3_decode.py, an
independent Layer III decoder written from the standard, to check that what the oracle emits decodes back to the input."""
import json, os, re, sys
src = open("/root/reference/src/main/java/mpg/TabInit.java").read()
body = re.search(r"dewin\[\]\s*=\s*\{(.*?)\};", src, re.S).group(1)
vals = [float(x) for x in re.findall(r"-?\d+\.\d+", body)]
assert len(vals) == 257, len(vals)
out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "synth_window.json")
json.dump({"source": "zhuker/lamejs src/main/java/mpg/TabInit.java dewin[] (ISO 11172-3 Table 3-B.3 D[i], i = 0..256, prototype signs)",
           "dewin": vals}, open(out, "w"))
print(out, len(vals), vals[255], vals[256])
