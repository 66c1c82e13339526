"""Import-path shim: the reference's Demo scripts import ``Models.fno.*`` (e.g.
/root/reference/Demo/Turbulence_3d+t/run_train_DDP.py:31) or, with ``Models/`` on sys.path,
``fno.*`` (e.g. /root/reference/Demo/Darcy_2d/run_FNO&UNet.py:17).  Both resolve to
deno4pytorch_b200."""
