"""ORACLE -- test infrastructure, not product code.

CPU restatement of the reference's per-slice segmentation path (see pipeline.py,
unet_torch.py, c/medseg_oracle.c).  Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs may import or execute anything under oracle/.
"""
