import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rd_vio_b200.frontend import FrontEnd
from conftest_shim import random_image
fe = FrontEnd(752, 480, 3, 21, num_slots=2, max_points=64)
a, b = fe.acquire(), fe.acquire()
img = random_image(480, 752, 1)
fe.preprocess([a, b], [img, img])
pts = np.array([[100.0, 100.0], [300.5, 200.25], [600.0, 400.0]])
print("tracking", flush=True)
n, s = fe.track([a], [b], [pts], None)
print(n, s)
