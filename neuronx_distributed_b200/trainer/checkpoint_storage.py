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

    def find_files(self, dirname: str, pattern: str) -> List[str]:
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

    def is_checkpoint_xser(self, dirname: str) -> bool:
        d = self._p(dirname)
        if not os.path.isdir(d):
            return False
        return any(x.endswith(".tensors") for x in os.listdir(d))

    def list_dirs(self, dirname: str) -> List[str]:
        d = self._p(dirname)
        if not os.path.isdir(d):
            return []
        return [x for x in os.listdir(d) if os.path.isdir(os.path.join(d, x))]

    def find_files(self, dirname, pattern=None, search_root: Optional[str] = None, max_count: Optional[int] = None,
                   sort_by_mdate: bool = False) -> List[str]:
        """Two call forms: ``find_files(dirname, pattern)`` (everything below ``dirname``) and the reference's
        ``find_files(pattern, search_depth, search_root=None, max_count=None, sort_by_mdate=False)`` (:176-205), which bounds
        the walk depth, optionally sorts newest first and truncates."""
        import fnmatch

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


class S3CheckpointStorage(BaseCheckpointStorage):
    """``s3://bucket/prefix`` storage (reference :236-605).  Requires boto3."""

    MAX_RETRY = 10

    def __init__(self, dirname: str):
        super().__init__(dirname)
        try:
            import boto3  # type: ignore
        except ImportError as e:  # pragma: no cover - boto3 is not in the offline image
            raise RuntimeError("S3CheckpointStorage needs boto3, which is not installed in this environment") from e
        assert dirname.startswith("s3://")
        rest = dirname[len("s3://"):]
        self.bucket, _, self.prefix = rest.partition("/")
        self.s3 = boto3.client("s3")

    def _key(self, name: str) -> str:
        return os.path.normpath(os.path.join(self.prefix, name)).lstrip("./")

    def _retry(self, fn: Callable, *a, **k):  # pragma: no cover - needs network
        wait = wait_decrementing_with_jitter(max_sleep=int(os.environ.get("WORLD_SIZE", "50000")) / 10000)
        for attempt in range(1, self.MAX_RETRY + 1):
            try:
                return fn(*a, **k)
            except Exception as e:  # noqa: BLE001
                if not is_slow_down_error(e) or attempt == self.MAX_RETRY:
                    raise
                time.sleep(wait(attempt))

    def dir_exists(self, dirname: str) -> bool:  # pragma: no cover
        r = self._retry(self.s3.list_objects_v2, Bucket=self.bucket, Prefix=self._key(dirname).rstrip("/") + "/", MaxKeys=1)
        return r.get("KeyCount", 0) > 0

    def file_exists(self, filename: str) -> bool:  # pragma: no cover
        try:
            self._retry(self.s3.head_object, Bucket=self.bucket, Key=self._key(filename))
            return True
        except Exception:  # noqa: BLE001
            return False

    def is_checkpoint_xser(self, dirname: str) -> bool:  # pragma: no cover
        r = self._retry(self.s3.list_objects_v2, Bucket=self.bucket, Prefix=self._key(dirname).rstrip("/") + "/")
        return any(".tensors/" in o["Key"] for o in r.get("Contents", []))

    def list_dirs(self, dirname: str) -> List[str]:  # pragma: no cover
        pfx = self._key(dirname).rstrip("/") + "/" if dirname not in (".", "") else (self.prefix.rstrip("/") + "/" if self.prefix else "")
        r = self._retry(self.s3.list_objects_v2, Bucket=self.bucket, Prefix=pfx, Delimiter="/")
        return [c["Prefix"][len(pfx):].rstrip("/") for c in r.get("CommonPrefixes", [])]

    def create_dir(self, dirname: str, exist_ok: bool = True) -> None:
        pass  # S3 has no directories

    def create_shared_dir(self, dirname: str, exist_ok: bool = True, process_group=None) -> None:
        pass

    def remove_dir(self, dirname: str) -> None:  # pragma: no cover
        pfx = self._key(dirname).rstrip("/") + "/"
        while True:
            r = self._retry(self.s3.list_objects_v2, Bucket=self.bucket, Prefix=pfx)
            objs = [{"Key": o["Key"]} for o in r.get("Contents", [])]
            if not objs:
                break
            self._retry(self.s3.delete_objects, Bucket=self.bucket, Delete={"Objects": objs})

    def remove_file(self, filename: str) -> None:  # pragma: no cover
        self._retry(self.s3.delete_object, Bucket=self.bucket, Key=self._key(filename))

    def save_text(self, text: str, filename: str) -> None:  # pragma: no cover
        self._retry(self.s3.put_object, Bucket=self.bucket, Key=self._key(filename), Body=text.encode())

    def save_object(self, obj: Any, filename: str) -> None:  # pragma: no cover
        buf = io.BytesIO()
        torch.save(obj, buf)
        buf.seek(0)
        self._retry(self.s3.upload_fileobj, buf, self.bucket, self._key(filename))

    def load_object(self, filename: str, map_location=None) -> Any:  # pragma: no cover
        buf = io.BytesIO()
        self._retry(self.s3.download_fileobj, self.bucket, self._key(filename), buf)
        buf.seek(0)
        return torch.load(buf, map_location=map_location, weights_only=False)


def create_checkpoint_storage(dirname: str) -> BaseCheckpointStorage:
    return S3CheckpointStorage(dirname) if dirname.startswith("s3://") else FilesysCheckpointStorage(dirname)
