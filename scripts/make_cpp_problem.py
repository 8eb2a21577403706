import struct, sys
from pathlib import Path
import numpy as np
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from oracle import gp_oracle as orc
X, y, Xs = orc.synthetic(400, 4, seed=31, m=25)
with open(sys.argv[1], "wb") as f:
    f.write(struct.pack("<qqq", X.shape[0], X.shape[1], Xs.shape[0]))
    f.write(X.tobytes()); f.write(y.tobytes()); f.write(Xs.tobytes())
