#!/usr/bin/env python3
"""Test infrastructure (run in the build container, where /root/reference exists): ships the decoded pixels of the smallest
bgdehaze fixture of the reference (modules/bgdehaze/img/PIS_T1A_259.jpg, decoded by cv2.imread like bgdehaze/main.py:16) as
tests/golden/k3_pis_full.npz so that SURVEY K3 - Background_light on a full-resolution fixture, first-index tie rule
(BGDehaze.py:14-31) - is checked on the GPU box against the values make_golden.py stored in kat.json."""
import json, os, sys
import numpy as np
import cv2

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import uwip_oracle as O

REF = "/root/reference/modules"
GOLD = os.path.join(ROOT, "tests", "golden")
name = "PIS_T1A_259"
img = cv2.imread(os.path.join(REF, "bgdehaze", "img", name + ".jpg"))
kat = json.load(open(os.path.join(GOLD, "kat.json")))
assert O.crc32(img) == kat["K3"][name]["decoded_crc"], "decoder drift: regenerate kat.json with make_golden.py"
B, idx = O.background_light(O.normalize_frame(img), 15, True)
assert [float(v) for v in B] == kat["K3"][name]["B_first_index"] and list(idx) == kat["K3"][name]["idx"]
np.savez_compressed(os.path.join(GOLD, "k3_pis_full.npz"), **{name + "/full": img})
print(img.shape, os.path.getsize(os.path.join(GOLD, "k3_pis_full.npz")))
