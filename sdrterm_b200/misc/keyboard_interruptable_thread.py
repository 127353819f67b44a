"""``KeyboardInterruptableThread`` (reference: src/misc/keyboard_interruptable_thread.py:24-50):
a thread whose uncaught exception -- ``KeyboardInterrupt`` included -- first runs a caller-supplied
shutdown hook.  The reference starts its reader and processor threads this way so that Ctrl-C in
any of them sets the halt condition (src/sdrterm.py:195-231); this build's CLI does the same for
its reader thread."""
from __future__ import annotations

import sys
import threading
from typing import Callable


class KeyboardInterruptableThread(threading.Thread):
    def __init__(self, func: Callable[[], None], target: Callable, group=None, name=None, args=(), daemon=None):
        if func is None:
            raise ValueError('func cannot be None')
        super().__init__(group=group, target=target, name=name, args=args, daemon=daemon)
        self._handleException = func
        # the hook is process-wide, as in the reference: the most recently created thread's hook serves
        threading.excepthook = self.handleException

    def handleException(self, e) -> None:
        """``threading.excepthook``: run the shutdown hook (its own failure is reported, never
        raised), then report what ended the thread -- an interrupt through the interpreter's
        default hook, anything else as a trace line."""
        from .general_util import tprint
        try:
            self._handleException()
        except Exception as ex:
            tprint(ex)
        except BaseException as ex:                      # e.g. the hook itself was interrupted
            sys.__excepthook__(type(ex), ex, ex.__traceback__)
        if issubclass(e.exc_type, KeyboardInterrupt):
            sys.__excepthook__(e.exc_type, e.exc_value, e.exc_traceback)
        else:
            tprint(e.exc_value)
