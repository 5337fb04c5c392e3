import numpy as np, scipy.ndimage as ndi, sys
sys.path.insert(0, __import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__)))))
import bench
X,dY,D=bench.make_inputs(0)
N=256
c=D.copy()
for a in range(1,4): c=ndi.spline_filter1d(c,3,axis=a,mode='mirror')
g=np.arange(N)*(4/(N-1))
grids=np.meshgrid(g,g,g,indexing='ij')
d=[ndi.map_coordinates(c[h],grids,order=3,mode='mirror',prefilter=False).astype(np.float32) for h in range(3)]
idx=np.meshgrid(np.arange(N),np.arange(N),np.arange(N),indexing='ij')
src=[idx[h]+d[h] for h in range(3)]
inr=np.ones((N,N,N),bool)
for h in range(3): inr&=(src[h]>=0)&(src[h]<=N-1)
st=[np.floor(src[h]).astype(np.int32)-1 for h in range(3)]
np.savez('/tmp/edf_boxes.npz',st0=st[0],st1=st[1],st2=st[2],inr=inr)
print('in range frac',inr.mean())
