"""Worker roster (torchlib/run_websocket_server.py:6-8 of the reference).

The reference's ``configs/websetting/config.csv`` holds one COLUMN per worker,

    id,alice,bob,charlie,crypto_provider
    host,127.0.0.1,...
    port,8777,...

read with ``pandas.read_csv(path, header=None, index_col=0).to_dict()`` -> ``{1: {"id": "alice", "host": .., "port": ..}, 2: ..}``
(keys = column numbers, values keyed by the first column).  This restatement returns the same structure (values as
strings or ints exactly as pandas infers them is not needed by any caller: ids are strings, ports are formatted with ``{:s}``
only in the websocket branch, which is out of scope) and also accepts the row-per-worker layout with an ``id,host,port``
header that round 1 of this repo used."""
import csv


def read_websocket_config(path: str):
    with open(path, newline="") as fh:
        rows = [r for r in csv.reader(fh) if r and any(c.strip() for c in r)]
    if not rows:
        return {}
    first_col = [r[0].strip() for r in rows]
    if first_col[:1] == ["id"] and "host" in first_col[1:]:          # the reference's layout: column per worker
        n = max(len(r) for r in rows)
        return {c: {r[0].strip(): r[c].strip() for r in rows if c < len(r)} for c in range(1, n)}
    header = [c.strip() for c in rows[0]]                             # row per worker with a header line
    if "id" not in header:
        raise ValueError(f"{path}: neither a column-per-worker roster (first column id/host/port) nor a table with an 'id' header")
    return {i: dict(zip(header, (c.strip() for c in r))) for i, r in enumerate(rows[1:], start=1)}


def worker_names(path: str):
    """ids in roster order -- ``[id_dict["id"] for _, id_dict in worker_dict.items()]`` (torchlib/utils.py:522)"""
    return [d["id"] for _, d in read_websocket_config(path).items()]
