"""Checkpoint storage back-ends (reference ``trainer/checkpoint_storage.py:46-612``).

``BaseCheckpointStorage`` is the small file-system API the checkpoint logic is written against;
``FilesysCheckpointStorage`` implements it for local/NFS/Lustre paths, ``S3CheckpointStorage`` for
``s3://`` URIs (boto3 is optional in this image — the class raises a clear error at construction if it
is missing).  Throttling errors are retried with decreasing jitter."""
from __future__ import annotations

import io
import os
import random
import shutil
import time
from abc import ABC, abstractmethod
from typing import Optional, Any, Callable, List

import torch


MB = 1024 ** 2
SHM_PATH = "/dev/shm"          # scratch for objects staged before an upload (reference checkpoint_storage.py:42-44)


class BaseCheckpointStorage(ABC):
    def __init__(self, dirname: str):
        self._dirname = dirname

    def dirname(self) -> str:
        return self._dirname

    # -- primitive ops ------------------------------------------------------------------
    @abstractmethod
    def dir_exists(self, dirname: str) -> bool: ...
    @abstractmethod
    def file_exists(self, filename: str) -> bool: ...
    @abstractmethod
    def is_checkpoint_xser(self, dirname: str) -> bool: ...
    @abstractmethod
    def list_dirs(self, dirname: str) -> List[str]: ...
    @abstractmethod
    def create_dir(self, dirname: str, exist_ok: bool = True) -> None: ...
    @abstractmethod
    def create_shared_dir(self, dirname: str, exist_ok: bool = True, process_group=None) -> None: ...
    @abstractmethod
    def remove_dir(self, dirname: str) -> None: ...
    @abstractmethod
    def remove_file(self, filename: str) -> None: ...
    @abstractmethod
    def save_text(self, text: str, filename: str) -> None: ...
    @abstractmethod
    def save_object(self, obj: Any, filename: str) -> None: ...
    @abstractmethod
    def load_object(self, filename: str, map_location=None) -> Any: ...

    # -- derived ops -------------------------------------------------------------------
    def is_checkpoint_tag_completed(self, tag: str) -> bool:
        return self.file_exists(os.path.join(tag, "done"))

    def list_checkpoint_tags(self) -> List[str]:
        """All tags (sub-directories containing a ``checkpoint`` marker), oldest first."""
        tags = [d for d in self.list_dirs(".") if self.file_exists(os.path.join(d, "checkpoint"))]
        return self._sort_tags(tags)

    def list_completed_checkpoint_tags(self) -> List[str]:
        return [t for t in self.list_checkpoint_tags() if self.is_checkpoint_tag_completed(t)]

    def find_files(self, dirname=None, pattern=None, search_root: Optional[str] = None, max_count: Optional[int] = None,
                   sort_by_mdate: bool = False, search_depth: Optional[int] = None) -> List[str]:
        return []

    def find_subdirs_contain_path(self, pattern: str, search_depth: int, search_root: Optional[str] = None,
                                  max_count: Optional[int] = None, sort_by_mdate: bool = False) -> List[str]:
        """Directories (relative to the storage root) that contain a file matching ``pattern`` at most ``search_depth``
        levels below ``search_root`` (reference :65-77) — e.g. every tag directory that has a ``done`` marker."""
        files = self.find_files(pattern, search_depth + 1, search_root, max_count, sort_by_mdate)
        return [os.path.dirname(f) for f in files]

    def remove_dirs(self, dirnames: List[str]) -> None:
        for d in dirnames:
            self.remove_dir(d)

    def remove_files(self, filenames: List[str]) -> None:
        for f in filenames:
            if self.file_exists(f):
                self.remove_file(f)

    def _sort_tags(self, tags: List[str]) -> List[str]:
        return sorted(tags, key=lambda t: self._tag_time(t))

    def _tag_time(self, tag: str) -> float:
        return 0.0


class FilesysCheckpointStorage(BaseCheckpointStorage):
    def _p(self, name: str) -> str:
        return os.path.join(self._dirname, name)

    def dir_exists(self, dirname: str) -> bool:
        return os.path.isdir(self._p(dirname))

    def file_exists(self, filename: str) -> bool:
        return os.path.isfile(self._p(filename))

    def is_checkpoint_xser(self, ckpt_path: str) -> bool:
        dirname = ckpt_path      # reference parameter names in the signature
        d = self._p(dirname)
        if not os.path.isdir(d):
            return False
        return any(x.endswith(".tensors") for x in os.listdir(d))

    def list_dirs(self, dirname: str) -> List[str]:
        d = self._p(dirname)
        if not os.path.isdir(d):
            return []
        return [x for x in os.listdir(d) if os.path.isdir(os.path.join(d, x))]

    def find_files(self, dirname=None, pattern=None, search_root: Optional[str] = None, max_count: Optional[int] = None,
                   sort_by_mdate: bool = False, search_depth: Optional[int] = None) -> List[str]:
        """Two call forms: ``find_files(dirname, pattern)`` (everything below ``dirname``) and the reference's
        ``find_files(pattern, search_depth, search_root=None, max_count=None, sort_by_mdate=False)`` (:176-205), which bounds
        the walk depth, optionally sorts newest first and truncates."""
        import fnmatch

        if search_depth is not None:                                   # keyword form of the reference's signature
            dirname, pattern = (pattern if dirname is None else dirname), search_depth
        if isinstance(pattern, int):
            pat, depth, root_rel = dirname, pattern, search_root or ""
        else:
            pat, depth, root_rel = pattern, None, dirname
        base = self._p(root_rel)
        out = []
        for root, dirs, files in os.walk(base):
            level = 0 if root == base else os.path.relpath(root, base).count(os.sep) + 1
            if depth is not None and level >= depth:
                dirs[:] = []
                if level > depth:
                    continue
            for f in files:
                if fnmatch.fnmatch(f, pat):
                    out.append(os.path.relpath(os.path.join(root, f), self._dirname))
        if sort_by_mdate:
            out.sort(key=lambda f: os.path.getmtime(self._p(f)), reverse=True)
        return out[:max_count] if max_count is not None else out

    def create_dir(self, dirname: str, exist_ok: bool = True) -> None:
        os.makedirs(self._p(dirname), exist_ok=exist_ok)

    def create_shared_dir(self, dirname: str, exist_ok: bool = True, process_group=None) -> None:
        import torch.distributed as dist

        rank = dist.get_rank(process_group) if dist.is_initialized() else 0
        if rank == 0:
            self.create_dir(dirname, exist_ok)
        if dist.is_initialized():
            dist.barrier(group=process_group)

    def remove_dir(self, dirname: str) -> None:
        shutil.rmtree(self._p(dirname), ignore_errors=True)

    def remove_file(self, filename: str) -> None:
        try:
            os.remove(self._p(filename))
        except FileNotFoundError:
            pass

    def save_text(self, text: str, filename: str) -> None:
        path = self._p(filename)
        os.makedirs(os.path.dirname(path), exist_ok=True)
        tmp = path + ".tmp"
        with open(tmp, "w") as f:
            f.write(text)
        os.replace(tmp, path)

    def save_object(self, obj: Any, filename: str) -> None:
        path = self._p(filename)
        if type(obj).__name__ == "_RawBytes":            # pre-serialised payload (trainer/checkpoint.py)
            os.makedirs(os.path.dirname(path), exist_ok=True)
            tmp = path + ".tmp"
            with open(tmp, "wb") as f:
                f.write(obj.data)
            os.replace(tmp, path)
            return
        os.makedirs(os.path.dirname(path), exist_ok=True)
        tmp = path + ".tmp"
        torch.save(obj, tmp)
        os.replace(tmp, path)   # atomic publish: a reader never sees a half-written file

    def load_object(self, filename: str, map_location=None) -> Any:
        return torch.load(self._p(filename), map_location=map_location, weights_only=False)

    def _tag_time(self, tag: str) -> float:
        try:
            return os.path.getmtime(self._p(os.path.join(tag, "checkpoint")))
        except OSError:
            return 0.0


def is_slow_down_error(exception: BaseException) -> bool:
    """S3 asks clients to back off with these error codes (reference :250-268); anything else is a real failure."""
    msg = str(exception)
    return any(f"<Code>{c}</Code>" in msg or c in msg for c in ("SlowDown", "RequestTimeout", "InternalError", "Throttling"))


class wait_decrementing_with_jitter:  # noqa: N801  (reference spelling, :236-241)
    """Back-off policy: sleep a random 1…⌈max_sleep / attempt⌉ seconds — the window SHRINKS with every attempt because by
    then the thundering herd of ranks has already been spread out by the first, widest draw."""

    def __init__(self, max_sleep: float) -> None:
        self.max_sleep = max_sleep

    def __call__(self, retry_state) -> float:
        import math

        attempt = getattr(retry_state, "attempt_number", retry_state)
        return float(random.randint(1, max(1, math.ceil(self.max_sleep / max(1, int(attempt))))))


_s3_resource = None
_s3_client = None
_s3_transfer_manager = None


def _retrying(fn: Callable) -> Callable:
    """Call ``fn`` again (≤ ``MAX_RETRY`` attempts, decreasing jitter) while S3 answers with a back-off error."""
    import functools

    @functools.wraps(fn)
    def wrapped(*a, **k):
        wait = wait_decrementing_with_jitter(max_sleep=int(os.environ.get("WORLD_SIZE", "50000")) / 10000)
        for attempt in range(1, S3CheckpointStorage.MAX_RETRY + 1):
            try:
                return fn(*a, **k)
            except Exception as e:  # noqa: BLE001
                if not is_slow_down_error(e) or attempt == S3CheckpointStorage.MAX_RETRY:
                    raise
                time.sleep(wait(attempt) * S3CheckpointStorage.SLEEP_SCALE)
    return wrapped


class S3CheckpointStorage(BaseCheckpointStorage):
    """``s3://bucket/prefix`` storage (reference :286-608).  Needs ``boto3`` (import-gated: not part of the offline image; the
    tests run this class against an in-memory stand-in of the client).  One client per process; every request goes through
    the throttling-aware retry."""

    S3_PATH_PREFIX = "s3://"
    MAX_RETRY = 10
    SLEEP_SCALE = 1.0                        # tests set this to 0 so that injected SlowDown errors do not sleep
    retry_with_jitter = staticmethod(_retrying)

    def __init__(self, dirname: str, crt_config: Optional[dict] = None):
        super().__init__(dirname)
        self._bucket, self._base_key = S3CheckpointStorage.parse_path(dirname)
        if self._base_key and not self._base_key.endswith("/"):
            self._base_key += "/"
        self.crt_config = dict(crt_config or {})
        S3CheckpointStorage.get_client()     # fail early (clear message) when boto3 is missing

    # ---- names ---------------------------------------------------------------------------------------------------------
    @staticmethod
    def parse_path(s3_path: str):
        """``"s3://bucket/a/b"`` → ``("bucket", "a/b")``; ``"s3://bucket"`` → ``("bucket", None)``."""
        head = S3CheckpointStorage.S3_PATH_PREFIX
        if not s3_path.startswith(head):
            raise RuntimeError(f"Error: invalid s3 path: {s3_path} because it does not start with {head}")
        rest = s3_path[len(head):]
        if not rest:
            raise RuntimeError(f"Error: invalid s3 path: {s3_path} that is empty")
        bucket, sep, key = rest.partition("/")
        return bucket, (key or None) if sep else None

    def convert_path_to_key(self, path: str) -> str:
        path = os.path.normpath(path).lstrip("./") if path not in ("", ".") else ""
        return path if self._base_key is None else self._base_key + path

    _key = convert_path_to_key

    @property
    def bucket(self) -> str:
        return self._bucket

    @property
    def prefix(self) -> str:
        return (self._base_key or "").rstrip("/")

    # ---- client --------------------------------------------------------------------------------------------------------
    @staticmethod
    def _ensure_s3_resource_and_client():
        global _s3_resource, _s3_client
        if _s3_client is None:
            try:
                import boto3  # type: ignore
            except ImportError as e:
                raise RuntimeError("S3CheckpointStorage needs boto3, which is not installed in this environment") from e
            try:
                import botocore.config  # type: ignore

                cfg = botocore.config.Config(max_pool_connections=max(1, (os.cpu_count() or 1) // 4))
                _s3_resource = boto3.Session().resource("s3", config=cfg)
                _s3_client = _s3_resource.meta.client
            except ImportError:
                _s3_client = boto3.client("s3")
                _s3_resource = None
        return _s3_resource, _s3_client

    @staticmethod
    def get_resource():
        return S3CheckpointStorage._ensure_s3_resource_and_client()[0]

    @staticmethod
    def get_client():
        return S3CheckpointStorage._ensure_s3_resource_and_client()[1]

    @property
    def s3(self):
        return S3CheckpointStorage.get_client()

    def get_transfer_manager(self, config=None):
        """Process-wide ``s3transfer`` manager for multi-part transfers (``None`` when s3transfer is unavailable: single
        requests are used then)."""
        global _s3_transfer_manager
        if _s3_transfer_manager is None:
            try:
                from boto3.s3.transfer import TransferConfig, create_transfer_manager  # type: ignore

                _s3_transfer_manager = create_transfer_manager(self.s3, config or TransferConfig())
            except Exception:  # noqa: BLE001
                _s3_transfer_manager = None
        return _s3_transfer_manager

    def _retry(self, fn: Callable, *a, **k):
        return _retrying(fn)(*a, **k)

    def _list(self, prefix: str, delimiter: Optional[str] = None, max_keys: Optional[int] = None):
        """All pages of ``list_objects_v2`` → (contents, common prefixes)."""
        contents, prefixes, token = [], [], None
        while True:
            kw = dict(Bucket=self._bucket, Prefix=prefix)
            if delimiter:
                kw["Delimiter"] = delimiter
            if max_keys:
                kw["MaxKeys"] = max_keys
            if token:
                kw["ContinuationToken"] = token
            r = self._retry(self.s3.list_objects_v2, **kw)
            contents += r.get("Contents", [])
            prefixes += [c["Prefix"] for c in r.get("CommonPrefixes", [])]
            token = r.get("NextContinuationToken")
            if not token or (max_keys and len(contents) + len(prefixes) >= max_keys):
                return contents, prefixes

    def _dir_key(self, dirname: str) -> str:
        k = self.convert_path_to_key(dirname)
        return k if (not k or k.endswith("/")) else k + "/"

    # ---- BaseCheckpointStorage -----------------------------------------------------------------------------------------
    def dir_exists(self, dirname: str) -> bool:
        contents, prefixes = self._list(self._dir_key(dirname), max_keys=1)
        return bool(contents or prefixes)

    def file_exists(self, filename: str) -> bool:
        key = self.convert_path_to_key(filename)
        contents, _ = self._list(key, max_keys=1)
        return any(o["Key"] == key for o in contents)

    def is_checkpoint_xser(self, dirname: str) -> bool:
        contents, _ = self._list(self._dir_key(dirname))
        return any(".tensors/" in o["Key"] for o in contents)

    def list_dirs(self, dirname: str) -> List[str]:
        pfx = self._dir_key(dirname)
        _, prefixes = self._list(pfx, delimiter="/")
        return [p[len(pfx):].rstrip("/") for p in prefixes]

    def find_files(self, dirname=None, pattern=None, search_root: Optional[str] = None, max_count: Optional[int] = None,
                   sort_by_mdate: bool = True, search_depth: Optional[int] = None) -> List[str]:
        """Two call forms, like the file-system back-end: ``find_files(dirname, pattern)`` — files under ``dirname`` whose
        base name matches the glob; ``find_files(pattern, search_depth, search_root, max_count, sort_by_mdate)`` — the
        reference's depth-limited search (paths relative to the storage root)."""
        import fnmatch

        if search_depth is not None:
            dirname, pattern = (pattern if dirname is None else dirname), search_depth
        if isinstance(pattern, int):
            glob_pat, depth, root = dirname, pattern, search_root or ""
        else:
            glob_pat, depth, root = pattern or "*", None, dirname
        pfx = self._dir_key(root)
        contents, _ = self._list(pfx)
        base = self._base_key or ""
        hits = []
        for o in contents:
            rel = o["Key"][len(pfx):]
            if depth is not None and rel.count("/") >= depth:
                continue
            if fnmatch.fnmatch(os.path.basename(rel), glob_pat):
                hits.append((o.get("LastModified", 0), o["Key"][len(base):]))
        if sort_by_mdate:
            hits.sort(key=lambda t: t[0])
        out = [h[1] for h in hits]
        return out[:max_count] if max_count else out

    def create_dir(self, dirname: str, exist_ok: bool = True) -> None:
        pass  # S3 has no directories

    def create_shared_dir(self, dirname: str, exist_ok: bool = True, process_group=None) -> None:
        pass

    def remove_dir(self, dirname: str) -> None:
        pfx = self._dir_key(dirname)
        while True:
            contents, _ = self._list(pfx, max_keys=1000)
            objs = [{"Key": o["Key"]} for o in contents[:1000]]
            if not objs:
                break
            self._retry(self.s3.delete_objects, Bucket=self._bucket, Delete={"Objects": objs})

    def remove_file(self, filename: str) -> None:
        self._retry(self.s3.delete_object, Bucket=self._bucket, Key=self.convert_path_to_key(filename))

    def save_text(self, text: str, filename: str, use_threads: bool = True) -> None:   # one PUT: nothing to thread
        self._retry(self.s3.put_object, Bucket=self._bucket, Key=self.convert_path_to_key(filename), Body=text.encode())

    def upload_stream_to_file(self, stream_creator, filename: str, chunk_size_MB: int = 64, max_concurrency: int = 10,
                              use_threads: bool = True) -> None:
        """Upload what ``stream_creator`` produces — an object with ``create_stream()`` (reference), a callable returning a
        binary stream, or the stream itself."""
        stream = stream_creator.create_stream() if hasattr(stream_creator, "create_stream") else (
            stream_creator() if callable(stream_creator) else stream_creator)
        stream.seek(0)
        kw = {}
        try:
            from boto3.s3.transfer import TransferConfig  # type: ignore

            kw["Config"] = TransferConfig(use_threads=use_threads, multipart_chunksize=chunk_size_MB << 20, max_concurrency=max_concurrency)
        except Exception:  # noqa: BLE001
            pass
        key = self.convert_path_to_key(filename)

        def attempt():
            stream.seek(0)                    # a failed attempt may have consumed part of the stream
            self.s3.upload_fileobj(stream, self._bucket, key, **kw)

        self._retry(attempt)

    def download_file_to_stream(self, filename: str, chunk_size_MB: int = 64, max_concurrency: int = 15) -> io.BytesIO:
        """One ``get_object`` (no extra transfer threads: tensor loading already runs from several threads per process)."""
        r = self._retry(self.s3.get_object, Bucket=self._bucket, Key=self.convert_path_to_key(filename))
        stream = io.BytesIO(r["Body"].read())
        stream.seek(0)
        return stream

    def save_object(self, obj: Any, filename: str) -> None:
        buf = io.BytesIO()
        if type(obj).__name__ == "_RawBytes":            # pre-serialised payload (trainer/checkpoint.py)
            buf.write(obj.data)
        else:
            torch.save(obj, buf)
        self.upload_stream_to_file(buf, filename)

    def load_object(self, filename: str, map_location=None) -> Any:
        return torch.load(self.download_file_to_stream(filename), map_location=map_location, weights_only=False)

    def load_text(self, filename: str) -> str:
        return self.download_file_to_stream(filename).read().decode()

    def _tag_time(self, tag: str) -> float:
        key = self.convert_path_to_key(os.path.join(tag, "checkpoint"))
        contents, _ = self._list(key, max_keys=1)
        for o in contents:
            if o["Key"] == key:
                lm = o.get("LastModified", 0)
                return lm.timestamp() if hasattr(lm, "timestamp") else float(lm)
        return 0.0


def create_checkpoint_storage(dirname: str, crt_config: Optional[dict] = None) -> BaseCheckpointStorage:
    return S3CheckpointStorage(dirname, crt_config) if dirname.startswith(S3CheckpointStorage.S3_PATH_PREFIX) \
        else FilesysCheckpointStorage(dirname)
