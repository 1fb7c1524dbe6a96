import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, cv2
import rm_radar_b200 as rr
from tests import fixtures as fx
exp = np.load(os.path.join(fx.GOLDEN, "expected_all.npz"))
img = cv2.imread(os.path.join(fx.GOLDEN, "_assets", "8.jpg"))
det = rr.RobotDetector(fx.engine("car"), fx.engine("armor"), fx.IMAGE_SIZE, fx.CLASS_NUM, fx.MAX_BATCH, fx.OPT_BATCH)
det.detect(img)
cars = [d.as_array() for d in det.last_cars()]
print("counts got", [len(det.last_armors(k)) for k in range(len(cars))], "want", exp["f8_armor_counts"].tolist())
for k in range(len(cars)):
    for d in det.last_armors(k):
        print(" got car", k, d.as_array())
print("want armors", exp["f8_armors"])
# armor head of car 4 vs the fp32 oracle, around the decision
if fx.have_onnx():
    from oracle import detect_oracle as do
    from oracle.onnx_torch import OnnxNet
    c = cars[4]
    roi = np.ascontiguousarray(img[int(c[1]):int(c[1]) + int(c[3]), int(c[0]):int(c[0]) + int(c[2])])
    a = rr.Detector(fx.engine("armor"), fx.CLASS_NUM, (roi.shape[1], roi.shape[0]), 1, conf_thresh=0.5)
    a.detect(roi)
    got = a.last_output(1)[0]
    x, _ = do.preprocess(roi)
    ref = OnnxNet(fx.onnx("armor"))(x[None]).numpy()[0]
    top = np.argsort(-ref[4:].max(axis=0))[:5]
    print("top anchors ref conf", ref[4:].max(axis=0)[top], "got", got[4:].max(axis=0)[top], "max |dconf|", np.abs(got[4:] - ref[4:]).max())
