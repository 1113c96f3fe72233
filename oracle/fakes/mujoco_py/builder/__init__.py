class _Cymj:
    @staticmethod
    def set_warning_callback(fn):
        pass


cymj = _Cymj()
