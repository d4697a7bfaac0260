"""In-memory-ish stand-in for the part of ``boto3``'s S3 client the checkpoint storage uses.  Objects live as files under
``$FAKE_S3_ROOT/<bucket>/`` (key url-quoted into one file name) so that several test processes see the same "bucket".
``$FAKE_S3_SLOWDOWN_EVERY=N`` makes every N-th request fail with an S3 ``SlowDown`` error to exercise the retry path."""
import io
import os
import urllib.parse

_calls = 0


class ClientError(Exception):
    pass


def _maybe_throttle():
    global _calls
    n = int(os.environ.get("FAKE_S3_SLOWDOWN_EVERY", "0"))
    _calls += 1
    if n and _calls % n == 0:
        raise ClientError("An error occurred (SlowDown) when calling the operation: <Error><Code>SlowDown</Code></Error>")


class _Client:
    def _dir(self, bucket):
        d = os.path.join(os.environ["FAKE_S3_ROOT"], bucket)
        os.makedirs(d, exist_ok=True)
        return d

    def _path(self, bucket, key):
        return os.path.join(self._dir(bucket), urllib.parse.quote(key, safe=""))

    def _keys(self, bucket):
        return sorted(urllib.parse.unquote(f) for f in os.listdir(self._dir(bucket)) if not f.endswith(".tmp~"))

    def put_object(self, Bucket, Key, Body):  # noqa: N803
        _maybe_throttle()
        tmp = self._path(Bucket, Key) + f".{os.getpid()}.tmp~"
        with open(tmp, "wb") as f:
            f.write(Body if isinstance(Body, bytes) else Body.read())
        os.replace(tmp, self._path(Bucket, Key))
        return {}

    def upload_fileobj(self, Fileobj, Bucket, Key, **kw):  # noqa: N803
        return self.put_object(Bucket, Key, Fileobj.read())

    def get_object(self, Bucket, Key):  # noqa: N803
        _maybe_throttle()
        try:
            with open(self._path(Bucket, Key), "rb") as f:
                return {"Body": io.BytesIO(f.read())}
        except FileNotFoundError:
            raise ClientError(f"NoSuchKey: {Key}") from None

    def download_fileobj(self, Bucket, Key, Fileobj, **kw):  # noqa: N803
        Fileobj.write(self.get_object(Bucket, Key)["Body"].read())

    def head_object(self, Bucket, Key):  # noqa: N803
        _maybe_throttle()
        if not os.path.exists(self._path(Bucket, Key)):
            raise ClientError("404 Not Found")
        return {"ContentLength": os.path.getsize(self._path(Bucket, Key))}

    def _delete(self, bucket, key):
        try:
            os.remove(self._path(bucket, key))
        except FileNotFoundError:
            pass

    def delete_object(self, Bucket, Key):  # noqa: N803
        _maybe_throttle()
        self._delete(Bucket, Key)
        return {}

    def delete_objects(self, Bucket, Delete):  # noqa: N803
        _maybe_throttle()                              # one request, however many keys
        for o in Delete["Objects"]:
            self._delete(Bucket, o["Key"])
        return {}

    def list_objects_v2(self, Bucket, Prefix="", Delimiter=None, MaxKeys=1000, ContinuationToken=None):  # noqa: N803
        _maybe_throttle()
        keys = [k for k in self._keys(Bucket) if k.startswith(Prefix)]
        contents, prefixes = [], []
        for k in keys:
            rest = k[len(Prefix):]
            if Delimiter and Delimiter in rest:
                p = Prefix + rest.split(Delimiter, 1)[0] + Delimiter
                if p not in prefixes:
                    prefixes.append(p)
            else:
                try:
                    contents.append({"Key": k, "LastModified": os.path.getmtime(self._path(Bucket, k)), "Size": os.path.getsize(self._path(Bucket, k))})
                except FileNotFoundError:
                    pass
        entries = [("c", c) for c in contents] + [("p", p) for p in prefixes]
        start = int(ContinuationToken or 0)
        page = entries[start:start + MaxKeys]
        out = {"Contents": [e for t, e in page if t == "c"], "CommonPrefixes": [{"Prefix": e} for t, e in page if t == "p"],
               "KeyCount": len(page), "IsTruncated": start + MaxKeys < len(entries)}
        if out["IsTruncated"]:
            out["NextContinuationToken"] = str(start + MaxKeys)
        return out


def client(name, **kw):
    assert name == "s3"
    return _Client()
