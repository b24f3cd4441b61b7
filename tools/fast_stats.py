"""Probe behind DESIGN.md §4: share of pixels that are FAST-9/16 corners at iniThFAST / minThFAST on a benchmark frame, and
how many pass the cheap rejection tests (4 compass points / 8 even ring positions) an early-out would use."""
import numpy as np, sys
import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from visual_sgraphs_b200.synth import synth_frame
import cv2
img = synth_frame(1000, 640, 480).astype(np.int32)
ring = [(0,3),(1,3),(2,2),(3,1),(3,0),(3,-1),(2,-2),(1,-3),(0,-3),(-1,-3),(-2,-2),(-3,-1),(-3,0),(-3,1),(-2,2),(-1,3)]
H,W = img.shape
c = img[3:H-3,3:W-3]
P = np.stack([img[3+dy:H-3+dy, 3+dx:W-3+dx] for dx,dy in ring])
for t in (20,7):
    br = P > c + t
    dk = P < c - t
    def arc9(m):
        out = np.zeros_like(m[0])
        for s in range(16):
            a = np.ones_like(m[0])
            for k in range(9):
                a &= m[(s+k)%16]
            out |= a
        return out
    corner = arc9(br) | arc9(dk)
    def compass(m):
        out = np.zeros_like(m[0])
        for a,b in ((0,4),(4,8),(8,12),(12,0)):
            out |= m[a] & m[b]
        return out
    comp = compass(br) | compass(dk)
    def even8(m):
        out = np.zeros_like(m[0])
        for s in range(0,16,2):
            a = np.ones_like(m[0])
            for k in range(0,8,2):
                a &= m[(s+k)%16]
            out |= a
        return out
    ev = even8(br) | even8(dk)
    print("t=%d corners %.3f compass-pass %.3f even8-pass %.3f" % (t, corner.mean(), comp.mean(), ev.mean()))
    # pair pass rate (two pixels S apart): approx independent
    print("  pair-pass compass %.3f even %.3f corner-pair %.3f" % (1-(1-comp.mean())**2, 1-(1-ev.mean())**2, 1-(1-corner.mean())**2))
