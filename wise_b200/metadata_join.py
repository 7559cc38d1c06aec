"""Batched post-search metadata join for WISE's CLI.

/root/reference/search.py:137-153 resolves every hit with two SELECTs (VectorRepo.get + MediaRepo.get); once the
search itself takes milliseconds that loop IS the query time at `--topk 1000` (docs/Retrieval-Evaluation.md:36-45:
0.307 s per query).  The REST API already batches (get_full_metadata_batch,
/root/reference/src/repository/__init__.py:42-79: one `IN (...)` query, results put back in rank order).  This is the
same join for the CLI, written against the DB-API so that it needs neither SQLAlchemy models nor pydantic:

    result = join_hits_batched(conn, ids, dist)      # conn: sqlite3.Connection or a SQLAlchemy Connection

Tables (/root/reference/src/db/tables/__init__.py:14-47): vectors(id, media_id, timestamp, end_timestamp),
media(id, path, ...).  Returns the dict process_text_query builds (search.py:154-159).
"""
from __future__ import annotations

_SQL = ("SELECT v.id, v.timestamp, v.end_timestamp, m.path FROM vectors v JOIN media m ON m.id = v.media_id "
        "WHERE v.id IN (%s)")
_CHUNK = 900  # SQLite's default limit is 999 bound variables per statement


def _rows(conn, sql, params):
    if hasattr(conn, "exec_driver_sql"):  # SQLAlchemy 2.x Connection
        return conn.exec_driver_sql(sql, tuple(params)).fetchall()
    return conn.execute(sql, tuple(params)).fetchall()  # sqlite3 / any DB-API connection with execute()


def join_hits_batched(conn, ids, dist):
    """ids / dist: one result row of index.search (ids == -1 marks the end, search.py:140-143)."""
    hit_ids = []
    for v in ids:
        v = int(v)
        if v == -1:
            break
        hit_ids.append(v)
    found = {}
    uniq = list(dict.fromkeys(hit_ids))
    for s in range(0, len(uniq), _CHUNK):
        part = uniq[s:s + _CHUNK]
        for vid, ts, end_ts, path in _rows(conn, _SQL % ",".join("?" * len(part)), part):
            found[int(vid)] = (ts, end_ts, path)
    missing = [v for v in uniq if v not in found]
    if missing:  # same failure as get_full_metadata_batch
        raise RuntimeError(f"Unable to retrieve metadata for all ids. Retrieved metadata for {len(found)}/{len(uniq)} ids")
    files, pts, scores = [], [], []
    for rank, vid in enumerate(hit_ids):
        ts, end_ts, path = found[vid]
        files.append(path)
        pts.append(ts if end_ts is None else [ts, end_ts])
        scores.append(float(dist[rank]))
    return {"match_filename_list": files, "match_pts_list": pts, "match_score_list": scores}
