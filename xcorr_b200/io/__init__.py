"""Host-side output of the cc variable set (mirror of ``xcorr.io``): SDS-like
paths, sha256 metadata / data hashes and the file writer."""
from .filesystem import ncfile  # noqa: F401
from .hasher import hash_dataset, hash_variable, to_json  # noqa: F401
from .write import write, read  # noqa: F401
