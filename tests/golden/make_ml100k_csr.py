"""Generates tests/golden/ml100k_csr.npz from the reference's in-tree fixture /root/reference/data.csv
(MovieLens-100K: header user_id,item_id,rating,timestamp; 100,000 rows), through the ORACLE's CSR builder
(stable sort by (user, timestamp), data.rs:236-265).  /root/reference does not exist on the GPU box, so the
derived arrays are committed:  user_ptr u32[944+1], item_ids u16[100000], timestamps u32[100000], plus the
raw triplets' input order (users u16, items u16, ts u32) so the sort itself can be re-checked.
Run:  python tests/golden/make_ml100k_csr.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import oracle_lib as O  # noqa: E402

d = np.loadtxt("/root/reference/data.csv", delimiter=",", skiprows=1)
users, items, ts = d[:, 0].astype(np.uint64), d[:, 1].astype(np.uint64), d[:, 3].astype(np.uint64)
nu, ni = int(users.max()) + 1, int(items.max()) + 1  # data.rs:202-203 (max + 1)
up, ii, tt = O.compress(users, items, ts, nu)
assert nu == 944 and ni == 1683 and len(ii) == 100000
np.savez_compressed(os.path.join(HERE, "ml100k_csr.npz"), user_ptr=up.astype(np.uint32), item_ids=ii.astype(np.uint16),
                    timestamps=tt.astype(np.uint32), raw_users=users.astype(np.uint16), raw_items=items.astype(np.uint16),
                    raw_ts=ts.astype(np.uint32), num_users=np.array(nu), num_items=np.array(ni))
print("wrote ml100k_csr.npz", os.path.getsize(os.path.join(HERE, "ml100k_csr.npz")), "bytes")
