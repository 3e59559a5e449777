#!/usr/bin/env python
"""BASELINE.json names ``train_federated.py``; in the reference ``train_federated`` is a function
(torchlib/utils.py:936) reached through ``train.py --train_federated``.  This shim forwards to the same main()."""
import sys

from train import main

if __name__ == "__main__":
    main(["--train_federated", "--unencrypted_aggregation"] + sys.argv[1:])
