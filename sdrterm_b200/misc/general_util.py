"""Host utilities of the reference's ``misc.general_util`` that the demodulation path's callers use
(reference src/misc/general_util.py:29-154): the three verbosity levels, exception printing,
socket shutdown, port lookup, and the signal handling that turns SIGINT/SIGTERM/... into the shared
``isDead`` flag (sdrterm.py:172-231).  No compute here."""
from __future__ import annotations

import os
import signal
import socket
import sys
import traceback
from typing import Callable

_LEVEL = 0          # 0 = errors only, 1 = -v, 2 = -vv


def eprint(*args, **kwargs) -> None:
    print(*args, file=sys.stderr, flush=True, **kwargs)


def vprint(*args, **kwargs) -> None:
    if _LEVEL >= 1:
        eprint(*args, **kwargs)


def tprint(*args, **kwargs) -> None:
    if _LEVEL >= 2:
        eprint(*args, **kwargs)


def verboseOn() -> None:
    global _LEVEL
    _LEVEL = max(_LEVEL, 1)


def traceOn() -> None:
    global _LEVEL
    _LEVEL = 2


def printException(e: BaseException, *_) -> None:
    eprint(f'Error: {e}')
    if _LEVEL >= 1:
        traceback.print_exception(type(e), e, e.__traceback__, file=sys.stderr)


def shutdownSocket(*socks: socket.socket) -> None:
    for s in socks:
        try:
            s.shutdown(socket.SHUT_RDWR)
        except OSError:
            pass


def findPort(host: str = 'localhost') -> int:
    with socket.socket(socket.AF_INET, socket.SOCK_STREAM) as s:
        s.bind((host, 0))
        return s.getsockname()[1]


def setSignalHandlers(pid: int, func: Callable[[], None]) -> list[int]:
    """SIGTERM, SIGINT (+ SIGQUIT, SIGABRT, SIGHUP, SIGXCPU on posix) print ``pid N caught: NAME``
    and call ``func`` (which sets the halt flag); keyboard backgrounding signals are ignored
    (general_util.py:136-154).  Must be called from the main thread."""
    sigs = [signal.SIGTERM, signal.SIGINT]
    if os.name == 'posix':
        sigs += [signal.SIGQUIT, signal.SIGABRT, signal.SIGHUP, signal.SIGXCPU]
        for s in (signal.SIGTSTP, signal.SIGTTIN, signal.SIGTTOU):
            signal.signal(s, lambda n, _f: vprint(f'Ignored signal {signal.Signals(n).name}'))

    def handle(n, _frame):
        eprint(f'pid {pid} caught: {signal.Signals(n).name}')
        func()

    for s in sigs:
        signal.signal(s, handle)
    return [int(s) for s in sigs]
