class Error(Exception):
    pass


class AlreadyPendingCallError(Error):
    pass


class NoAsyncCallError(Error):
    pass
