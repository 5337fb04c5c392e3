import numpy as np
z=np.load('/tmp/edf_boxes.npz'); st=[z['st0'],z['st1'],z['st2']]; inr=z['inr']
N=256; NT=4
def chunk_stats(G,MR,TX, rows_cap=364, maxq=16):
    # chunk = G z-slabs x MR rows x TX x
    shp=(N//G,G,N//MR,MR,N//TX,TX)
    big=10**6
    mn=[];mx=[]
    for h in range(3):
        a=st[h].reshape(shp)
        m=inr.reshape(shp)
        mn.append(np.where(m,a,big).min(axis=(1,3,5)))
        mx.append(np.where(m,a,-big).max(axis=(1,3,5)))
    empty=mn[0]>mx[0]
    nzw=mx[0]-mn[0]+NT; nyw=mx[1]-mn[1]+NT
    wx0=mn[2]&~3; nq=((mx[2]+NT-1-wx0)>>2)+1
    rows=nzw*nyw
    fit=(~empty)&(nq<=maxq)&(rows<=rows_cap)
    ne=(~empty).sum()
    return dict(chunks=int(ne), fit=float(fit.sum()/ne), fail_q=float(((~empty)&(nq>maxq)).sum()/ne), fail_rows=float(((~empty)&(rows>rows_cap)).sum()/ne),
                rows_med=float(np.median(rows[~empty])), rows_p95=float(np.percentile(rows[~empty],95)), rows_p99=float(np.percentile(rows[~empty],99)), nq_med=float(np.median(nq[~empty])), nq_p99=float(np.percentile(nq[~empty],99)),
                cells_per_voxel=float((rows[fit]*nq[fit]*4).sum()/ (fit.sum()*G*MR*TX)))
print('8x4x32 cap364', chunk_stats(8,4,32))
print('8x4x32 cap480/q16', chunk_stats(8,4,32,480))
print('8x4x32 cap728 (pitch 32?)', chunk_stats(8,4,32,728,8))
print('4x4x32 cap364', chunk_stats(4,4,32))
print('8x2x32 cap364', chunk_stats(8,2,32))
print('8x8x16', chunk_stats(8,8,16))
print('--- capacity sweep for the bench field (8x4x32 chunks)')
for cap in (208, 240, 280, 320, 364, 420):
    r=chunk_stats(8,4,32,cap); print(cap, 'fit %.4f'%r['fit'])
print('16 z-slabs x 4 x 32 (512-thread CTA):', {k:round(v,3) for k,v in chunk_stats(16,4,32,364).items()})
