def get_trigger(trigger):
    return trigger
