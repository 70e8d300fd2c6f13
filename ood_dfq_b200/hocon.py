"""Minimal reader for the reference's ``config/*.hocon`` files.

The reference parses them with pyhocon (``options.py:12``), which is not installed here.  Its files only use a
flat subset of HOCON -- ``key = value`` lines, ``#`` / ``//`` comments, numbers, quoted or bare strings,
booleans and one-line lists -- which is all this reader accepts (anything else raises, loudly).

``load(path)`` returns a dict; ``QuantSettings.from_file(path)`` picks out what drives the quantisation path and
applies the reference's hard-coded overrides (``lam`` and ``eps`` are fixed in ``options.py:64-65`` whatever the
file says).
"""
from __future__ import annotations

import re
from dataclasses import dataclass

_NUM = re.compile(r"^[+-]?(\d+\.?\d*([eE][+-]?\d+)?|\.\d+([eE][+-]?\d+)?)$")


def _strip_comment(line: str) -> str:
    out, quote = [], None
    i = 0
    while i < len(line):
        ch = line[i]
        if quote:
            out.append(ch)
            if ch == quote:
                quote = None
        elif ch in "\"'":
            quote = ch
            out.append(ch)
        elif ch == "#" or line.startswith("//", i):
            break
        else:
            out.append(ch)
        i += 1
    return "".join(out).strip()


def _scalar(tok: str):
    tok = tok.strip()
    if len(tok) >= 2 and tok[0] == tok[-1] and tok[0] in "\"'":
        return tok[1:-1]
    low = tok.lower()
    if low in ("true", "yes", "on"):
        return True
    if low in ("false", "no", "off"):
        return False
    if low == "null":
        return None
    if _NUM.match(tok):
        return int(tok) if re.match(r"^[+-]?\d+$", tok) else float(tok)
    return tok                                  # bare string (e.g. model_name = resnet18)


def _value(text: str):
    text = text.strip().rstrip(",")
    if text.startswith("["):
        if not text.endswith("]"):
            raise ValueError(f"unterminated list: {text!r} (multi-line lists are not supported)")
        inner = text[1:-1].strip()
        return [] if not inner else [_scalar(t) for t in inner.split(",")]
    if text.startswith("{"):
        raise ValueError("nested objects are not supported by this reader")
    return _scalar(text)


def loads(text: str) -> dict:
    conf = {}
    for ln, raw in enumerate(text.splitlines(), 1):
        line = _strip_comment(raw)
        if not line:
            continue
        m = re.match(r"^([A-Za-z_][\w.\-]*)\s*[=:]\s*(.*)$", line)
        if not m:
            raise ValueError(f"line {ln}: cannot parse {raw!r}")
        conf[m.group(1)] = _value(m.group(2))
    return conf


def load(path: str) -> dict:
    with open(path) as f:
        return loads(f.read())


@dataclass
class QuantSettings:
    """The options that reach the quantisation path (``options.py:12-71``)."""
    model_name: str
    dataset: str
    batchSize: int
    nClasses: int
    img_size: int
    channels: int
    qw: int
    qa: int
    temperature: float
    alpha: float
    lr_S: float
    momentum: float
    weightDecay: float
    lam: float = 1000.0          # options.py:64: hard-coded, the file's value is ignored
    eps: float = 0.01            # options.py:65: hard-coded

    @classmethod
    def from_dict(cls, c: dict) -> "QuantSettings":
        return cls(model_name=str(c["model_name"]), dataset=str(c["dataset"]), batchSize=int(c["batchSize"]),
                   nClasses=int(c["nClasses"]), img_size=int(c["img_size"]), channels=int(c["channels"]),
                   qw=int(c["qw"]), qa=int(c["qa"]), temperature=float(c["temperature"]), alpha=float(c["alpha"]),
                   lr_S=float(c["lr_S"]), momentum=float(c["momentum"]), weightDecay=float(c["weightDecay"]))

    @classmethod
    def from_file(cls, path: str) -> "QuantSettings":
        return cls.from_dict(load(path))
