"""Generates tests/golden/maildir_small_tfidf_sample.npz from the reference's own corpus
(/root/reference/data/maildir_small, 8586 Enron mails): the whole corpus goes through the ETL
restatement (apss_b200.etl), IDF is fitted on all documents, and a fixed sample of the documents
(every 6th file of the sorted listing; duplicated folders such as sent / sent_items stay together because
the sample is by position) is stored un-normalised.  Run in the build container only."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import apss_b200
from apss_b200 import etl

root = sys.argv[1] if len(sys.argv) > 1 else "/root/reference/data/maildir_small"
t0 = time.time()
paths = etl.list_files(root)
indptr, indices, values, idf, m = etl.tfidf_corpus(paths)
print("corpus: %d docs, %d components, %.1fs" % (m, len(indices), time.time() - t0))
# sample: whole folders that are known duplicates of each other + a stride sample of the rest
sel = [i for i, p in enumerate(paths) if ("/arora-h/" in p) or (i % 9 == 0)]
ip = [0]
ix, vv = [], []
for i in sel:
    a, b = indptr[i], indptr[i + 1]
    ix.append(indices[a:b]); vv.append(values[a:b]); ip.append(ip[-1] + (b - a))
rel = [os.path.relpath(paths[i], root) for i in sel]
out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "maildir_small_tfidf_sample.npz")
np.savez_compressed(out, indptr=np.asarray(ip, np.int64), indices=np.concatenate(ix).astype(np.int32),
                    values=np.concatenate(vv).astype(np.float64), paths=np.array(rel), n_docs_corpus=m,
                    df_null=int(round((m + 1) / np.exp(idf[etl.HashingTF().index_of("null")]) - 1)))
print("wrote %s: %d docs, %d components, %.1f KB" % (out, len(sel), ip[-1], os.path.getsize(out) / 1e3))
