def set_session(session):
    return None
