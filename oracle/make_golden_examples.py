"""Fixtures for BASELINE.json configs[0]: the raw LSD rows of the reference's four bundled example images
(assets/examples/*.jpg), produced with the reference's own detector.

The detector (lsdpython/lsd_1.6/lsd.c, upstream of the hot path) is compiled from where it lies under
/root/reference into oracle/_ref/liblsd.so (git-ignored) and called with the defaults of the reference's
Cython wrapper (lsdpython/lsd.pyx:17-18).  Pre-processing follows evaluation.py:139-154 with the
substitutions SURVEY.md section 8(c) lists: cv2.resize (area interpolation) for ImageMagick's
`convert -resize 640x640`, rgb2gray = 0.2125 R + 0.7154 G + 0.0721 B (skimage 0.12).
Output: tests/golden/examples_lsd.npz (rows (N,7) in pixels + image shapes).
  python oracle/make_golden_examples.py
"""
import ctypes as C
import glob
import os
import subprocess

import cv2
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"
LIB = os.path.join(HERE, "_ref", "liblsd.so")
OUT = os.path.join(HERE, "..", "tests", "golden", "examples_lsd.npz")


def build_lsd():
    os.makedirs(os.path.dirname(LIB), exist_ok=True)
    subprocess.check_call(["gcc", "-O2", "-shared", "-fPIC", "-o", LIB, os.path.join(REF, "lsdpython", "lsd_1.6", "lsd.c"), "-lm"])
    lib = C.CDLL(LIB)
    lib.LineSegmentDetection.restype = C.POINTER(C.c_double)
    lib.LineSegmentDetection.argtypes = [C.POINTER(C.c_int), C.c_void_p, C.c_int, C.c_int, C.c_double, C.c_double, C.c_double,
                                         C.c_double, C.c_double, C.c_double, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
    return lib


def detect(lib, image):
    """lsd.pyx:17-43 with its default arguments."""
    img = np.ascontiguousarray(image, np.float64)
    n = C.c_int(0)
    p = lib.LineSegmentDetection(C.byref(n), img.ctypes.data, img.shape[1], img.shape[0], 0.8, 0.6, 2.0, 22.5, 0.0, 0.7, 1024,
                                 None, None, None)
    return np.ctypeslib.as_array(p, shape=(n.value, 7)).copy()


def main():
    lib = build_lsd()
    g = {}
    files = sorted(glob.glob(os.path.join(REF, "assets", "examples", "*.jpg")))
    for i, f in enumerate(files):
        bgr = cv2.imread(f)
        h, w = bgr.shape[:2]
        s = 640.0 / max(w, h)                                   # convert -resize 640x640: fit inside, keep the aspect
        if s < 1.0:
            bgr = cv2.resize(bgr, (int(round(w * s)), int(round(h * s))), interpolation=cv2.INTER_AREA)
        rgb = bgr[:, :, ::-1].astype(np.float64) / 255.0
        gray = 0.2125 * rgb[:, :, 0] + 0.7154 * rgb[:, :, 1] + 0.0721 * rgb[:, :, 2]     # color.rgb2gray (evaluation.py:150)
        image = gray.astype("float64")                          # detect_lsd_lines (evaluation.py:229-231)
        if np.max(image) <= 1:
            image = image * 255
        rows = detect(lib, image)
        g["lsd_%d" % i] = rows
        g["shape_%d" % i] = np.array(image.shape)
        g["name_%d" % i] = np.array(os.path.basename(f))
        print(os.path.basename(f), image.shape, rows.shape[0], "segments")
    g["n_images"] = np.array(len(files))
    np.savez_compressed(OUT, **g)
    print("wrote", OUT, os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    main()
