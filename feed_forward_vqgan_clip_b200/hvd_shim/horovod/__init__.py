"""Drop-in stand-in for the `horovod` package (see horovod/torch.py).  Put `feed_forward_vqgan_clip_b200/hvd_shim` on
sys.path (or call feed_forward_vqgan_clip_b200.parallel.install_horovod_shim()) and the reference's
`import horovod.torch as hvd` (main.py:45) resolves here, with the collectives running over NCCL / NVLink."""
