import sys
sys.path.insert(0, '.')
from visual_sgraphs_b200.extractor import ORBextractor
from visual_sgraphs_b200.synth import synth_frame
ex = ORBextractor(1000, max_batch=1)
f = synth_frame(1000, 640, 480)
ex(f); ex(f)
