import sys, subprocess, os
sys.path.insert(0,'.')
CASES = [  # C,h,w,s,K,N,frac
 (1,13,11,2,5,3,True), (1,13,11,2,5,3,False), (1,40,40,2,5,3,True), (1,64,64,2,5,3,True), (1,13,12,2,5,3,True),
 (1,14,11,2,5,3,True), (1,8,8,1,3,2,True), (2,33,17,2,0,4,True), (2,12,20,4,9,6,True), (1,40,40,4,7,16,False),
 (1,13,11,2,3,3,True), (1,13,11,2,7,3,True), (1,100,100,2,5,3,True),
]
if len(sys.argv) > 1:
    import numpy as np, srb200 as srb
    from importlib import import_module
    wl = import_module("super-resolution_b200.workloads")
    C,h,w,s,K,N,frac = eval(sys.argv[1])
    rng = np.random.default_rng(15)
    psf = wl.gaussian_psf(K, 1.0 + 0.25*K) if K else None
    shifts = rng.uniform(-2.5,2.5,size=(N,2)) if frac else rng.integers(-3,4,size=(N,2)).astype(float)
    x = rng.random((C,h*s,w*s)); lr = rng.random((N,C,h,w))
    from oracle import sr_oracle as o
    m = o.Model(s, psf, shifts); obs = o.upsample_observations(m, lr)
    fo, go = o.evaluate(m, x, obs, 0, 0.0, None)
    with srb.Engine(lr.shape, s, psf, shifts) as e:
        e.set_observations(lr)
        f,g = e.eval(x)
        print("path", e.active_path, "cost rel", abs(f-fo)/abs(fo), "grad rel", np.linalg.norm(g-go)/np.linalg.norm(go))
else:
    for c in CASES:
        r = subprocess.run([sys.executable, __file__, repr(c)], capture_output=True, text=True)
        print(c, r.stdout.strip()[-120:], (r.stderr.strip().splitlines() or [""])[-1][-150:])
