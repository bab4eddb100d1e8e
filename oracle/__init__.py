"""oracle/: CPU restatement of the reference's DINOv2 hot path -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / `--impl reference` legs may import this
package, and only as the checker / the CPU baseline. The product (lightly_train_b200) never imports it.
"""
