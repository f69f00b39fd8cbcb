"""pygame.sprite stand-in: insertion-ordered Group, Sprite membership, rect collisions.

TEST INFRASTRUCTURE ONLY.  See the package docstring.
"""


class Sprite:
    def __init__(self, *groups):
        # pygame re-initialises membership on every __init__ call; Robot.on_reset
        # (RR_Robot.py:87) calls __init__ again on a live sprite, which in real pygame
        # leaves the Groups still holding the sprite.  Keep that: only reset our view.
        self.__g = {}
        for g in groups:
            self.add(g)

    def add(self, *groups):
        for g in groups:
            if g not in self.__g:
                g.add_internal(self)
                self.add_internal(g)

    def remove(self, *groups):
        for g in groups:
            if g in self.__g:
                g.remove_internal(self)
                self.remove_internal(g)

    def add_internal(self, group):
        self.__g[group] = 0

    def remove_internal(self, group):
        self.__g.pop(group, None)

    def kill(self):
        for g in list(self.__g):
            g.remove_internal(self)
        self.__g.clear()

    def alive(self):
        return bool(self.__g)

    def groups(self):
        return list(self.__g)


class Group:
    def __init__(self, *sprites):
        self.spritedict = {}
        self.add(*sprites)

    def sprites(self):
        return list(self.spritedict)

    def add_internal(self, sprite):
        self.spritedict[sprite] = 0

    def remove_internal(self, sprite):
        self.spritedict.pop(sprite, None)

    def has_internal(self, sprite):
        return sprite in self.spritedict

    def add(self, *sprites):
        for s in sprites:
            if isinstance(s, Sprite):
                if not self.has_internal(s):
                    self.add_internal(s)
                    s.add_internal(self)
            else:
                self.add(*list(s))

    def remove(self, *sprites):
        for s in sprites:
            if isinstance(s, Sprite):
                if self.has_internal(s):
                    self.remove_internal(s)
                    s.remove_internal(self)
            else:
                self.remove(*list(s))

    def has(self, *sprites):
        return all(self.has_internal(s) for s in sprites)

    def empty(self):
        for s in self.sprites():
            self.remove(s)

    def __iter__(self):
        return iter(self.sprites())

    def __contains__(self, s):
        return self.has_internal(s)

    def __len__(self):
        return len(self.spritedict)

    def __bool__(self):
        return len(self.spritedict) > 0


def collide_rect(left, right):
    return left.rect.colliderect(right.rect)


def spritecollide(sprite, group, dokill, collided=None):
    fn = collided or collide_rect
    hit = [s for s in group.sprites() if fn(sprite, s)]
    if dokill:
        for s in hit:
            s.kill()
    return hit
