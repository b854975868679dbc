import sys, subprocess, os
sys.path.insert(0,'.')
CASES = [  # C,h,w,s,K,N,frac
 (1,13,11,2,5,3,True), (1,13,11,2,5,3,False), (1,160,160,2,5,3,True), (2,64,80,4,7,16,False), (1,96,128,3,7,9,False),
 (1,300,260,1,3,2,True), (2,130,170,2,0,4,True), (2,70,90,4,9,6,True), (1,100,100,4,7,16,False),
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
    wts = 0.5 + rng.random(x.shape)
    from oracle import sr_oracle as o
    m = o.Model(s, psf, shifts); obs = o.upsample_observations(m, lr)
    fo, go = o.evaluate(m, x, obs, 0, 0.02, wts, threads=8)
    with srb.Engine(lr.shape, s, psf, shifts) as e:
        e.set_observations(lr); e.set_regularizer(0, 0.02); e.set_irls_weights(wts)
        f,g = e.eval(x)
        print("path", e.active_path, "cost rel %.2e grad rel %.2e max %.2e" % (abs(f-fo)/abs(fo), np.linalg.norm(g-go)/np.linalg.norm(go), np.abs(g-go).max()/np.abs(go).max()))
else:
    for c in CASES:
        r = subprocess.run([sys.executable, __file__, repr(c)], capture_output=True, text=True)
        print(c, r.stdout.strip()[-120:], (r.stderr.strip().splitlines() or [""])[-1][-150:])
