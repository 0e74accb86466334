"""Import stub so /root/reference/common/db.py constructs without sqlalchemy (TEST INFRASTRUCTURE ONLY)."""


class _Any:
    def __init__(self, *a, **k):
        pass

    def __call__(self, *a, **k):
        return _Any()

    def __getattr__(self, name):
        return _Any()


Column = Integer = Float = String = ForeignKey = Table = MetaData = _Any


def create_engine(*a, **k):
    return _Any()
