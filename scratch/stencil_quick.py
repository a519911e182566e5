import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
exec(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'stencil_bulk_time.py')).read().split("def setenv")[0])
for (B, n) in ((1920, 64), (256, 256), (128, 64)):
    run('bulk', B, n)
