import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
from oracle.binding import Oracle
from idp_b200 import ContactContext
o=Oracle()
rng = np.random.default_rng(2)
Xs = rng.uniform(-1, 1, (40, 3))
Fs = np.array([rng.choice(39, 3, replace=False) for _ in range(120)], np.int32)
Fs[5] = Fs[4][[1, 0, 2]]
Fs[7, :2] = Fs[4, :2]
Xs[Fs[9, 2]] = Xs[Fs[9, 1]]
want=o.surface(40,Fs,Xs)
c=ContactContext(0)
c.set_mesh_from_triangles(40,Fs,Xs)
got=c.get_surface_primitives(areas=True)
for k in ("BNArea","BEArea","BTArea"):
    bad=np.nonzero(~np.isclose(got[k],want[k],rtol=1e-13,atol=1e-300))[0]
    print(k,bad, got[k][bad], want[k][bad], want['bedge'][bad] if k=="BEArea" else "")
print(Fs[4],Fs[5],Fs[7])
