"""Load tests/golden/*.npz (made by tests/golden/make_golden.py from the reference)."""
import glob
import hashlib
import os

import numpy as np

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
NAMES = sorted(os.path.splitext(os.path.basename(p))[0]
               for p in glob.glob(os.path.join(GOLDEN_DIR, "*.npz")))


class Golden:
    def __init__(self, name):
        z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
        self.name = name
        self.args = str(z["args"])
        self.seed = int(z["seed"])
        self.stages = int(z["stages"])
        (self.nx, self.ny, self.nz, self.num_vars, self.comm_vars, self.max_blocks,
         self.stencil, self.permute) = (int(x) for x in z["params"])
        self.slots, self.level = z["slots"], z["level"]
        self.nei_level, self.nei = z["nei_level"], z["nei"]
        self.check_sums = z["check_sums"]
        self.sha256 = str(z["sha256"])
        self.tile_shape = (self.nx + 2, self.ny + 2, self.nz + 2)
        # --stencil 0 fixtures: the coefficients init() drew (init.c:418-423)
        self.stencil0 = (int(z["s0_mat"]), float(z["s0_a1"]), np.asarray(z["s0_a0"], np.float64)) \
            if "s0_mat" in z.files else None

    def apply_stencil0(self, mesh):
        """hand the --stencil 0 coefficients to a DeviceMesh / OracleMesh (no-op for 7/27)"""
        if self.stencil0:
            mesh.set_stencil0(*self.stencil0)

    def seeded_blocks(self):
        """Yield (slot, tiles[num_vars, nx+2, ny+2, nz+2]) exactly as make_golden seeded them."""
        rs = np.random.RandomState(self.seed)
        shape = (self.num_vars,) + self.tile_shape
        for s in self.slots:
            yield int(s), rs.random_sample(shape)


def digest(blocks):
    h = hashlib.sha256()
    for b in blocks:
        h.update(np.ascontiguousarray(b, np.float64).tobytes())
    return h.hexdigest()
