import os, subprocess, sys
N = sys.argv[1] if len(sys.argv) > 1 else "262144"
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
vdir = os.path.join(root, "jitterbug_b200", "_lib", "variants")
for v in sorted(os.listdir(vdir)):
    env = dict(os.environ, JB_LIB=os.path.join(vdir, v, "libjbsim.so"))
    r = subprocess.run([sys.executable, os.path.join(root, "tests", "gpu_smoke1.py"), N], env=env, capture_output=True, text=True, timeout=120)
    print(v, '\n   '.join(r.stdout.strip().splitlines()[-2:]) if r.stdout.strip() else 'FAILED: ' + r.stderr.strip().splitlines()[-1][:200])
