import sys; sys.path.insert(0,'.')
import numpy as np
from imageanalysis_b200 import _capi
from oracle import orb as O
g=np.load('tests/golden/orb_reference.npz'); img=g['texture_img']
eng=_capi.Engine(_capi.NORM_HAMMING,32,0)
s=eng.debug_orb_fast(img).astype(int)
o=O.fast_scores(img)
print('score maps equal:', (s==o).all(), 'mismatches', int((s!=o).sum()), 'of', s.size)
ys,xs=np.nonzero(s!=o)
for y,x in list(zip(ys,xs))[:6]: print((x,y),'gpu',s[y,x],'oracle',o[y,x])
print('gpu row 3  :', s[3, :24].tolist())
print('oracle row3:', o[3, :24].tolist())
print('gpu col 3  :', s[:24, 3].tolist())
best = None
for dy in range(-3, 4):
    for dx in range(-3, 4):
        sh = np.roll(np.roll(o, dy, 0), dx, 1)
        m = int((sh[8:-8, 8:-8] != s[8:-8, 8:-8]).sum())
        if best is None or m < best[0]: best = (m, dx, dy)
print('best shift of the oracle map onto the gpu map (mismatches, dx, dy):', best)
ot = O.fast_scores(np.ascontiguousarray(img.T)).T
print('transposed-image hypothesis mismatches:', int((ot != s).sum()))
for t in (10, 15, 25, 30):
    print('threshold', t, 'mismatches', int((O.fast_scores(img, t) != s).sum()))
